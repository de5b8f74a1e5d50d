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


def _attr(name, value):
    """AttributeProto: float (f), int (i), str (s) or list of ints (ints), with its type tag (field 20)."""
    msg = _ld(1, name.encode())
    if isinstance(value, float):
        return msg + _vi((2 << 3) | 5) + struct.pack("<f", value) + _vi((20 << 3) | 0) + _vi(1)
    if isinstance(value, int):
        return msg + _vi((3 << 3) | 0) + _vi(value) + _vi((20 << 3) | 0) + _vi(2)
    if isinstance(value, str):
        return msg + _ld(4, value.encode()) + _vi((20 << 3) | 0) + _vi(3)
    return msg + b"".join(_vi((8 << 3) | 0) + _vi(v) for v in value) + _vi((20 << 3) | 0) + _vi(7)


def _node_proto(op, ins, outs, **attrs):
    msg = b"".join(_ld(1, i.encode()) for i in ins) + b"".join(_ld(2, o.encode()) for o in outs) + _ld(4, op.encode())
    return msg + b"".join(_ld(5, _attr(k, v)) for k, v in attrs.items())


def write_onnx(path, state, folded, final_sigmoid=True, resize_mode="nearest", swap_concat=False):
    """The AttU_Net graph as torch exports it in eval mode: convs in forward order with their
    attributes, MaxPool / Resize / Add / Relu / Sigmoid / Mul / Concat wiring; BatchNormalization nodes when not folded.
    The keyword switches produce the malformed variants the reader must refuse."""
    from stroke_derenderer_b200.weights import conv_bn_slots, fold_conv_bn
    slots = {conv: (n, bn, k) for n, (conv, bn, cout, cin, k) in enumerate(conv_bn_slots())}
    nodes, inits, uid = [], [], [0]

    def fresh(tag):
        uid[0] += 1
        return f"{tag}{uid[0]}"

    def conv(name, x):
        n, bn, k = slots[name]
        if folded:
            w, b = fold_conv_bn(state, name, bn)
        else:
            w, b = state[f"{name}.weight"], state[f"{name}.bias"]
        inits.extend([_tensor_proto(f"w{n}", w, raw=n % 2 == 0), _tensor_proto(f"b{n}", b)])
        out = fresh("c")
        nodes.append(_node_proto("Conv", [x, f"w{n}", f"b{n}"], [out], kernel_shape=[k, k], pads=[k // 2] * 4, strides=[1, 1],
                                 dilations=[1, 1], group=1))
        if bn is not None and not folded:
            for tag, key in (("g", "weight"), ("be", "bias"), ("m", "running_mean"), ("v", "running_var")):
                inits.append(_tensor_proto(f"{tag}{n}", state[f"{bn}.{key}"]))
            o2 = fresh("n")
            nodes.append(_node_proto("BatchNormalization", [out, f"g{n}", f"be{n}", f"m{n}", f"v{n}"], [o2], epsilon=1e-5))
            out = o2
        return out

    def unary(op, x, **attrs):
        out = fresh(op[0].lower())
        nodes.append(_node_proto(op, [x], [out], **attrs))
        return out

    def block(name, x):
        return unary("Relu", conv(f"{name}.conv.3", unary("Relu", conv(f"{name}.conv.0", x))))

    xs = [block("Conv1", "input")]
    for lvl in range(2, 6):
        xs.append(block(f"Conv{lvl}", unary("MaxPool", xs[-1], kernel_shape=[2, 2], strides=[2, 2])))
    d = xs[4]
    for lvl in range(5, 1, -1):
        up = fresh("u")
        nodes.append(_node_proto("Resize", [d, "", "scales"], [up], mode=resize_mode))
        d = unary("Relu", conv(f"Up{lvl}.up.1", up))
        skip = xs[lvl - 2]
        g1, x1 = conv(f"Att{lvl}.W_g.0", d), conv(f"Att{lvl}.W_x.0", skip)
        add = fresh("a")
        nodes.append(_node_proto("Add", [g1, x1], [add]))
        psi = unary("Sigmoid", conv(f"Att{lvl}.psi.0", unary("Relu", add)))
        mul = fresh("m")
        nodes.append(_node_proto("Mul", [skip, psi], [mul]))
        cat = fresh("k")
        nodes.append(_node_proto("Concat", [d, mul] if swap_concat else [mul, d], [cat], axis=1))
        d = block(f"Up_conv{lvl}", cat)
    out = conv("Conv_1x1", d)
    if final_sigmoid:
        out = unary("Sigmoid", out)
    graph = b"".join(_ld(1, nd) for nd in nodes) + _ld(2, b"torch_jit") + b"".join(_ld(5, t) for t in inits)
    model = _vi((1 << 3) | 0) + _vi(8) + _ld(2, b"pytorch") + _ld(7, graph)
    Path(path).write_bytes(model)
