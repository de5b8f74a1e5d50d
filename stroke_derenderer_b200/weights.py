"""Weights of the skeletonization Attention-UNet binarizer: layer table, the
seeded "parity" initialisation, BN folding and file I/O.

The reference ships no weights and no model definition (the `.onnx` lives on a
Drive link, /root/reference/README.md:27, :54); the topology is the upstream
`AttU_Net(img_ch=3, output_ch=1)` the README names. The
parameter names below follow that module's `state_dict()` so a real checkpoint
maps one to one.

This module is product code (the engine folds and packs from it); the oracle
(`oracle/attunet_torch.py`) consumes the same dictionaries.
"""

from __future__ import annotations

import numpy as np

BN_EPS = 1e-5

# (name, kind, cin, cout).  kind: "block" = conv_block (two conv3x3+BN+ReLU),
# "up" = up_conv (nearest x2 + conv3x3+BN+ReLU), "att" = attention gate
# (cin = F_g = F_l, cout = F_int), "head" = final 1x1 conv.
def layer_table(img_ch: int = 3, output_ch: int = 1, base: int = 64):
    c1, c2, c3, c4, c5 = base, base * 2, base * 4, base * 8, base * 16
    return [
        ("Conv1", "block", img_ch, c1),
        ("Conv2", "block", c1, c2),
        ("Conv3", "block", c2, c3),
        ("Conv4", "block", c3, c4),
        ("Conv5", "block", c4, c5),
        ("Up5", "up", c5, c4),
        ("Att5", "att", c4, c4 // 2),
        ("Up_conv5", "block", c5, c4),
        ("Up4", "up", c4, c3),
        ("Att4", "att", c3, c3 // 2),
        ("Up_conv4", "block", c4, c3),
        ("Up3", "up", c3, c2),
        ("Att3", "att", c2, c2 // 2),
        ("Up_conv3", "block", c3, c2),
        ("Up2", "up", c2, c1),
        ("Att2", "att", c1, c1 // 2),
        ("Up_conv2", "block", c2, c1),
        ("Conv_1x1", "head", c1, output_ch),
    ]


def conv_bn_slots(img_ch: int = 3, output_ch: int = 1, base: int = 64):
    """Every (conv_prefix, bn_prefix|None, cout, cin, k) in the network, in
    state_dict order."""
    slots = []
    for name, kind, cin, cout in layer_table(img_ch, output_ch, base):
        if kind == "block":
            slots.append((f"{name}.conv.0", f"{name}.conv.1", cout, cin, 3))
            slots.append((f"{name}.conv.3", f"{name}.conv.4", cout, cout, 3))
        elif kind == "up":
            slots.append((f"{name}.up.1", f"{name}.up.2", cout, cin, 3))
        elif kind == "att":
            slots.append((f"{name}.W_g.0", f"{name}.W_g.1", cout, cin, 1))
            slots.append((f"{name}.W_x.0", f"{name}.W_x.1", cout, cin, 1))
            slots.append((f"{name}.psi.0", f"{name}.psi.1", 1, cout, 1))
        elif kind == "head":
            slots.append((name, None, cout, cin, 1))
    return slots


def make_parity_weights(seed: int = 123, img_ch: int = 3, output_ch: int = 1,
                        base: int = 64) -> dict[str, np.ndarray]:
    """Seeded, well-conditioned random weights.

    PyTorch's default init gives logits with std 0.002 (all-ones mask), which
    is useless as a parity test; this recipe gives logits with std ~2.
    conv weight ~ N(0, 2/fan_in), bias ~ N(0, 0.05^2); BN gamma ~ U(0.5, 1.5),
    beta ~ N(0, 0.2^2), running_mean ~ N(0, 0.2^2), running_var ~ U(0.5, 1.5).
    """
    rng = np.random.default_rng(seed)
    sd: dict[str, np.ndarray] = {}
    for conv, bn, cout, cin, k in conv_bn_slots(img_ch, output_ch, base):
        fan_in = cin * k * k
        sd[f"{conv}.weight"] = (rng.standard_normal((cout, cin, k, k)) *
                                np.sqrt(2.0 / fan_in)).astype(np.float32)
        sd[f"{conv}.bias"] = (rng.standard_normal(cout) * 0.05).astype(np.float32)
        if bn is not None:
            sd[f"{bn}.weight"] = rng.uniform(0.5, 1.5, cout).astype(np.float32)
            sd[f"{bn}.bias"] = (rng.standard_normal(cout) * 0.2).astype(np.float32)
            sd[f"{bn}.running_mean"] = (rng.standard_normal(cout) * 0.2).astype(np.float32)
            sd[f"{bn}.running_var"] = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    return sd


def fold_conv_bn(sd: dict, conv: str, bn: str | None):
    """Eval-mode BN folded into the preceding conv, in float64 then fp32:
    s = gamma / sqrt(var + eps); w' = w * s; b' = (b - mean) * s + beta."""
    w = sd[f"{conv}.weight"].astype(np.float64)
    b = sd[f"{conv}.bias"].astype(np.float64)
    if bn is not None and f"{bn}.weight" in sd:       # a dict without BN entries is already folded (onnx_reader)
        s = sd[f"{bn}.weight"].astype(np.float64) / np.sqrt(
            sd[f"{bn}.running_var"].astype(np.float64) + BN_EPS)
        w = w * s[:, None, None, None]
        b = (b - sd[f"{bn}.running_mean"].astype(np.float64)) * s + sd[f"{bn}.bias"].astype(np.float64)
    return w.astype(np.float32), b.astype(np.float32)


def save_weights(path: str, sd: dict) -> None:
    np.savez(path, **sd)


def load_weights(path: str) -> dict[str, np.ndarray]:
    """Loads a checkpoint: a `.npz` state dict with the upstream parameter names, or an exported
    `binarizer.onnx` (/root/reference/main.py:43) through the dependency-free reader."""
    if str(path).lower().endswith(".onnx"):
        from .onnx_reader import load_onnx_state
        return load_onnx_state(str(path))
    with np.load(path) as z:
        return {k: z[k] for k in z.files}
