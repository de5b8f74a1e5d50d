"""CPU emulation of the GPU numerics (BN folded in fp32, fp16 operands, fp32
accumulate, one rounding per stored tensor) to find which layers' weight /
activation rounding dominates the logit error.  Test/dev tool (imports oracle/)."""
import sys, json, time
from pathlib import Path
import numpy as np, torch, torch.nn.functional as F
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
from stroke_derenderer_b200.weights import make_parity_weights, fold_conv_bn
from stroke_derenderer_b200.engine import _slot_sources
from stroke_derenderer_b200.synth import synth_line
from oracle import segmentation_ref as O

torch.set_grad_enabled(False)
h = lambda t: t.half().float()

def forward(x, W, rw, ra):
    """W: slot -> (w,b) fp32 tensors; rw(slot, w) rounds weights; ra(name, t) rounds activations."""
    def conv(slot, t, relu=True, pad=1):
        w, b = W[slot]; y = F.conv2d(t, rw(slot, w), b, padding=pad)
        return ra(slot, F.relu(y) if relu else y)
    def up(slot, t): return conv(slot, F.interpolate(t, scale_factor=2, mode="nearest"))
    def gate(n, g, xx):
        wg, bg = W[f"ATT{n}_G"]; wx, bx = W[f"ATT{n}_X"]; wp, bp = W[f"ATT{n}_PSI"]
        q = F.relu(F.conv2d(g, rw(f"ATT{n}_G", wg), bg) + F.conv2d(xx, rw(f"ATT{n}_X", wx), bx))
        psi = torch.sigmoid(F.conv2d(q, wp, bp))
        return ra(f"ATT{n}", xx * psi)
    x = ra("in", x)
    x1 = conv("CONV1_1", conv("CONV1_0", x)); x2 = conv("CONV2_1", conv("CONV2_0", F.max_pool2d(x1, 2)))
    x3 = conv("CONV3_1", conv("CONV3_0", F.max_pool2d(x2, 2))); x4 = conv("CONV4_1", conv("CONV4_0", F.max_pool2d(x3, 2)))
    x5 = conv("CONV5_1", conv("CONV5_0", F.max_pool2d(x4, 2)))
    d = x5
    for n, skip in [(5, x4), (4, x3), (3, x2), (2, x1)]:
        u = up(f"UP{n}", d); a = gate(n, u, skip)
        d = conv(f"UPCONV{n}_1", conv(f"UPCONV{n}_0", torch.cat((a, u), 1)))
    w, b = W["HEAD"]
    return F.conv2d(d, w, b)

def main():
    gold = json.loads((ROOT / "tests/golden/golden.json").read_text())
    st = make_parity_weights(123); st["Conv_1x1.bias"] = np.array([gold["unet"]["head_bias"]], np.float32)
    W = {s: tuple(torch.from_numpy(a) for a in fold_conv_bn(st, *src)) for s, src in _slot_sources().items()}
    stack, *_ = O.cut_and_stack([synth_line(1000, 9)], (1, 3, 128, 384), 64)
    x = torch.from_numpy((stack / 255.).astype(np.float32))
    ident_w = lambda s, w: w; ident_a = lambda s, t: t
    ref = forward(x, W, ident_w, ident_a)
    def report(tag, z):
        e = (z - ref).abs(); mm = ((z > 0) != (ref > 0)).float().mean().item() * 100
        print(f"{tag:34s} logit err mean {e.mean():.5f} max {e.max():.4f}  mask mismatch {mm:.4f}%", flush=True)
        return e.mean().item()
    report("all fp16 (weights+acts)", forward(x, W, lambda s, w: h(w), lambda s, t: h(t)))
    report("weights fp16 only", forward(x, W, lambda s, w: h(w), ident_a))
    report("acts fp16 only", forward(x, W, ident_w, lambda s, t: h(t)))
    slots = [s for s in W if not s.endswith("PSI") and s != "HEAD"]
    res = {}
    for s in slots:
        res[s] = report(f"  only W[{s}] fp16", forward(x, W, lambda k, w, s=s: h(w) if k == s else w, ident_a))
    tot = np.sqrt(sum(v * v for v in res.values()))
    print("quadrature sum of per-layer weight contributions:", tot)
    for s, v in sorted(res.items(), key=lambda kv: -kv[1]):
        print(f"   {s:12s} {v:.5f}  share of variance {v * v / tot ** 2 * 100:5.1f}%")
    acts = ["in"] + [s for s in slots if "ATT" not in s] + [f"ATT{n}" for n in (5, 4, 3, 2)]
    resa = {}
    for s in acts:
        resa[s] = report(f"  only act[{s}] fp16", forward(x, W, ident_w, lambda k, t, s=s: h(t) if k == s else t))
    tot = np.sqrt(sum(v * v for v in resa.values()))
    for s, v in sorted(resa.items(), key=lambda kv: -kv[1]):
        print(f"   act {s:12s} {v:.5f}  share of variance {v * v / tot ** 2 * 100:5.1f}%")

if __name__ == "__main__":
    main()
