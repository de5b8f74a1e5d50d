"""Minimal ONNX (protobuf wire format) writer, just enough to exercise `stroke_derenderer_b200.onnx_reader`:
a ModelProto whose GraphProto holds Conv (+ optional BatchNormalization) and Relu nodes in the forward order of
AttU_Net and the initializers they reference.  Test infrastructure only."""
import struct
from pathlib import Path

import numpy as np

def _vi(n):
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def _ld(fno, payload):
    return _vi((fno << 3) | 2) + _vi(len(payload)) + payload


def _tensor_proto(name, arr, raw=True):
    arr = np.ascontiguousarray(arr, np.float32)
    msg = b"".join(_vi((1 << 3) | 0) + _vi(d) for d in arr.shape) + _vi((2 << 3) | 0) + _vi(1)
    msg += _ld(8, name.encode())
    msg += _ld(9, arr.tobytes()) if raw else _ld(4, arr.tobytes())       # raw_data | packed float_data
    return msg


def _node_proto(op, ins, outs, eps=None):
    msg = b"".join(_ld(1, i.encode()) for i in ins) + b"".join(_ld(2, o.encode()) for o in outs) + _ld(4, op.encode())
    if eps is not None:
        msg += _ld(5, _ld(1, b"epsilon") + _vi((2 << 3) | 5) + struct.pack("<f", eps) + _vi((20 << 3) | 0) + _vi(1))
    return msg


def write_onnx(path, state, folded):
    from stroke_derenderer_b200.weights import conv_bn_slots, fold_conv_bn
    nodes, inits, cur = [], [], "input"
    for n, (conv, bn, cout, cin, k) in enumerate(conv_bn_slots()):
        if folded:
            w, b = fold_conv_bn(state, conv, bn)
        else:
            w, b = state[f"{conv}.weight"], state[f"{conv}.bias"]
        inits += [_tensor_proto(f"w{n}", w, raw=n % 2 == 0), _tensor_proto(f"b{n}", b)]
        nodes.append(_node_proto("Conv", [cur, f"w{n}", f"b{n}"], [f"c{n}"]))
        cur = f"c{n}"
        if bn is not None and not folded:
            for tag, key in (("g", "weight"), ("be", "bias"), ("m", "running_mean"), ("v", "running_var")):
                inits.append(_tensor_proto(f"{tag}{n}", state[f"{bn}.{key}"]))
            nodes.append(_node_proto("BatchNormalization", [cur, f"g{n}", f"be{n}", f"m{n}", f"v{n}"], [f"n{n}"], eps=1e-5))
            cur = f"n{n}"
        nodes.append(_node_proto("Relu", [cur], [f"r{n}"]))
        cur = f"r{n}"
    graph = b"".join(_ld(1, nd) for nd in nodes) + _ld(2, b"torch_jit") + b"".join(_ld(5, t) for t in inits)
    model = _vi((1 << 3) | 0) + _vi(8) + _ld(2, b"pytorch") + _ld(7, graph)
    Path(path).write_bytes(model)


