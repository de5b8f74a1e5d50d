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
    inputs, outputs, op, eps, attrs = [], [], "", BN_EPS, {}
    for fno, wt, val in _fields(buf):
        if fno == 1:
            inputs.append(bytes(val).decode())
        elif fno == 2:
            outputs.append(bytes(val).decode())
        elif fno == 4:
            op = bytes(val).decode()
        elif fno == 5:                              # AttributeProto: name (1), f (2), i (3), s (4), ints (8)
            aname, aval, ints = "", None, []
            for f2, w2, v2 in _fields(val):
                if f2 == 1:
                    aname = bytes(v2).decode()
                elif f2 == 2 and w2 == 5:
                    aval = struct.unpack("<f", bytes(v2))[0]
                elif f2 == 3 and w2 == 0:
                    aval = v2
                elif f2 == 4 and w2 == 2:
                    aval = bytes(v2).decode(errors="replace")
                elif f2 == 8:
                    if w2 == 0:
                        ints.append(v2)
                    else:                           # packed
                        p = 0
                        while p < len(v2):
                            d, p = _varint(v2, p)
                            ints.append(d)
            attrs[aname] = ints if ints else aval
            if aname == "epsilon" and isinstance(aval, float):
                eps = float(aval)
    return {"op": op, "in": inputs, "out": outputs, "eps": eps, "attrs": attrs}


def check_topology(nodes, path: str = "model") -> None:
    """The engine hard-codes the Attention-UNet data flow: a graph that merely has the right
    conv shapes but another wiring would load cleanly and binarize wrongly, so refuse it here.  Checked: the graph
    ENDS in a Sigmoid (evaluate_binarize.py:103 compares the output with bin_thr as a probability; the upstream
    AttU_Net returns logits unless the export adds the activation), 4 MaxPool, 4 nearest Resize / Upsample, 4 Concat
    whose FIRST input is the gated skip (a Mul) and whose second is the up-convolution path, 5 Sigmoid (4 psi + output),
    and every Conv is stride 1, group 1, dilation 1, 'same' padding."""
    def fail(msg):
        raise ValueError(f"{path}: not the Attention-UNet binarizer graph this engine implements: {msg}")
    producer = {o: nd for nd in nodes for o in nd["out"]}
    compute = [nd for nd in nodes if nd["op"] not in ("Constant", "Identity", "Shape", "Gather", "Unsqueeze", "Cast", "Slice")]
    if not compute or compute[-1]["op"] != "Sigmoid":
        fail(f"the graph ends in {compute[-1]['op'] if compute else 'nothing'}, not Sigmoid (the engine thresholds probabilities; "
             "re-export with the final sigmoid)")
    count = lambda op: sum(1 for nd in nodes if nd["op"] == op)
    ups = [nd for nd in nodes if nd["op"] in ("Resize", "Upsample")]
    for op, want in (("MaxPool", 4), ("Concat", 4), ("Sigmoid", 5), ("Mul", 4)):
        if count(op) != want:
            fail(f"{count(op)} {op} nodes, expected {want}")
    if len(ups) != 4:
        fail(f"{len(ups)} Resize/Upsample nodes, expected 4")
    for nd in ups:
        mode = nd["attrs"].get("mode", "nearest")
        if mode != "nearest":
            fail(f"{nd['op']} mode {mode!r}, the engine upsamples with nearest")
    for nd in nodes:
        if nd["op"] == "Concat":
            first = producer.get(nd["in"][0])
            if len(nd["in"]) != 2 or first is None or first["op"] != "Mul":
                fail("a Concat whose first input is not the gated skip tensor (x * psi)")
            if nd["attrs"].get("axis", 1) != 1:
                fail("a Concat that is not along the channel axis")
        if nd["op"] == "MaxPool" and (nd["attrs"].get("kernel_shape", [2, 2]) != [2, 2] or nd["attrs"].get("strides", [2, 2]) != [2, 2]):
            fail("a MaxPool that is not 2x2 / stride 2")
        if nd["op"] == "Conv":
            a = nd["attrs"]
            k = a.get("kernel_shape", [None])[0]
            if any(v != 1 for v in a.get("strides", [1, 1])) or any(v != 1 for v in a.get("dilations", [1, 1])) or a.get("group", 1) != 1:
                fail("a Conv with stride / dilation / group other than 1")
            if k is not None and any(v != k // 2 for v in a.get("pads", [k // 2] * 4)):
                fail(f"a {k}x{k} Conv whose padding is not {k // 2}")


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


def load_onnx_state(path: str, img_ch: int = 3, output_ch: int = 1, base: int = 64, check: bool = True) -> dict:
    """Folded conv weights of an exported AttU_Net as a state dict `<conv>.weight` / `<conv>.bias`.  `check`: also
    verify the wiring the engine hard-codes (`check_topology`)."""
    nodes, inits = read_graph(path)
    if check:
        check_topology(nodes, path)
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
