"""Dependency-free reader of the weights of an exported `binarizer.onnx`
(/root/reference/main.py:43,48 loads it with onnxruntime; README.md:27 puts the file on a Drive link).

Neither `onnx` nor `onnxruntime` can be installed offline, so this walks the protobuf wire format directly:
ModelProto.graph (field 7) -> GraphProto.node (1) / .initializer (5) -> TensorProto dims (1) / data_type (2) /
float_data (4) / name (8) / raw_data (9).  The Attention-UNet's convolutions are taken **in graph order**, which
for a torch export is the forward order of `AttU_Net` = the slot order of `weights.conv_bn_slots()`:
Conv1.0, Conv1.3, ..., Up5, Att5.W_g, Att5.W_x, Att5.psi, Up_conv5.0, ... Conv_1x1.
Eval-mode exports fold BatchNorm into the convs; if a `BatchNormalization` node still follows a conv it is folded
here (float64, like `weights.fold_conv_bn`).  The result is a state dict holding only `<conv>.weight` / `.bias`
(already folded), which `UNetEngine` accepts.

No real checkpoint is reachable offline: the reader is tested against files produced by a minimal writer of the
same wire format (tests/test_host.py), i.e. the parse and the mapping are verified, the Drive file itself is not.
"""

from __future__ import annotations

import struct

import numpy as np

from .weights import BN_EPS, conv_bn_slots

_DTYPES = {1: np.float32, 10: np.float16, 11: np.float64, 6: np.int32, 7: np.int64}


def _varint(buf: memoryview, pos: int):
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _fields(buf: memoryview):
    """Yields (field number, wire type, value) of one message; length-delimited values are memoryviews."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val, pos = buf[pos:pos + 8], pos + 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val, pos = buf[pos:pos + ln], pos + ln
        elif wt == 5:
            val, pos = buf[pos:pos + 4], pos + 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield fno, wt, val


def _tensor(buf: memoryview):
    dims, dtype, name, raw, floats = [], 1, "", None, []
    for fno, wt, val in _fields(buf):
        if fno == 1:
            if wt == 0:
                dims.append(val)
            else:                                   # packed repeated int64
                p = 0
                while p < len(val):
                    d, p = _varint(val, p)
                    dims.append(d)
        elif fno == 2:
            dtype = val
        elif fno == 4:
            if wt == 2:
                floats.append(np.frombuffer(val, dtype="<f4"))
            else:
                floats.append(np.frombuffer(val, dtype="<f4", count=1))
        elif fno == 8:
            name = bytes(val).decode()
        elif fno == 9:
            raw = val
        elif fno == 13 and len(val):
            raise ValueError(f"initializer {name!r} uses external data, which this reader does not follow")
    if dtype not in _DTYPES:
        return name, None
    if raw is not None:
        arr = np.frombuffer(raw, dtype=np.dtype(_DTYPES[dtype]).newbyteorder("<"))
    elif floats:
        arr = np.concatenate(floats)
    else:
        arr = np.zeros(0, _DTYPES[dtype])
    return name, arr.reshape(dims) if dims else arr


def _node(buf: memoryview):
    inputs, outputs, op, eps = [], [], "", BN_EPS
    for fno, wt, val in _fields(buf):
        if fno == 1:
            inputs.append(bytes(val).decode())
        elif fno == 2:
            outputs.append(bytes(val).decode())
        elif fno == 4:
            op = bytes(val).decode()
        elif fno == 5:                              # AttributeProto: name (1), f (2)
            aname, af = "", None
            for f2, w2, v2 in _fields(val):
                if f2 == 1:
                    aname = bytes(v2).decode()
                elif f2 == 2 and w2 == 5:
                    af = struct.unpack("<f", bytes(v2))[0]
            if aname == "epsilon" and af is not None:
                eps = float(af)
    return {"op": op, "in": inputs, "out": outputs, "eps": eps}


def read_graph(path: str):
    """-> (nodes in graph order, {initializer name: ndarray})."""
    data = memoryview(open(path, "rb").read())
    graph = None
    for fno, wt, val in _fields(data):
        if fno == 7 and wt == 2:
            graph = val
    if graph is None:
        raise ValueError(f"{path}: no GraphProto found (not an ONNX model?)")
    nodes, inits = [], {}
    for fno, wt, val in _fields(graph):
        if fno == 1 and wt == 2:
            nodes.append(_node(val))
        elif fno == 5 and wt == 2:
            name, arr = _tensor(val)
            if arr is not None:
                inits[name] = arr
    return nodes, inits


def load_onnx_state(path: str, img_ch: int = 3, output_ch: int = 1, base: int = 64) -> dict:
    """Folded conv weights of an exported AttU_Net as a state dict `<conv>.weight` / `<conv>.bias`."""
    nodes, inits = read_graph(path)
    consumers = {}
    for nd in nodes:
        for i in nd["in"]:
            consumers.setdefault(i, []).append(nd)
    convs = [nd for nd in nodes if nd["op"] == "Conv"]
    slots = conv_bn_slots(img_ch, output_ch, base)
    if len(convs) != len(slots):
        raise ValueError(f"{path}: {len(convs)} Conv nodes, the Attention-UNet binarizer has {len(slots)}")
    sd = {}
    for nd, (conv, _bn, cout, cin, k) in zip(convs, slots):
        if len(nd["in"]) < 2 or nd["in"][1] not in inits:
            raise ValueError(f"{path}: weights of Conv '{conv}' are not an initializer")
        w = np.asarray(inits[nd["in"][1]], np.float64)
        if w.shape != (cout, cin, k, k):
            raise ValueError(f"{path}: Conv '{conv}' has weights {w.shape}, expected {(cout, cin, k, k)}")
        b = np.asarray(inits[nd["in"][2]], np.float64) if len(nd["in"]) > 2 and nd["in"][2] in inits else np.zeros(cout)
        nxt = consumers.get(nd["out"][0], [])
        if len(nxt) == 1 and nxt[0]["op"] == "BatchNormalization":       # export without BN folding
            bn = nxt[0]
            g, beta, mean, var = (np.asarray(inits[n], np.float64) for n in bn["in"][1:5])
            s = g / np.sqrt(var + bn["eps"])
            w = w * s[:, None, None, None]
            b = (b - mean) * s + beta
        sd[f"{conv}.weight"] = w.astype(np.float32)
        sd[f"{conv}.bias"] = b.astype(np.float32)
    return sd
