"""Debug: run sd_ccl_label repeatedly on the same masks and report any difference between runs / vs cv2."""
import sys, time
from pathlib import Path
import numpy as np, torch, cv2
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
from stroke_derenderer_b200 import segment as S
from stroke_derenderer_b200.synth import config_widths, ink_mask, synth_line
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
t0 = time.time()
widths = config_widths(512)[:n]
masks = [ink_mask(synth_line(int(w), seed=i)) for i, w in enumerate(widths)]
print("synth", round(time.time() - t0, 1), "s")
dev = torch.device("cuda", 0)
batch = S.plan_batch([m.shape[1] for m in masks], dev)
host = np.zeros(batch.px_total, np.uint8)
for m, ln in zip(masks, batch.lines):
    off, pitch = int(ln["px_off"]), int(ln["pitch"])
    host[off:off + 128 * pitch].reshape(128, pitch)[:, :m.shape[1]] = m
planes = torch.from_numpy(host).to(dev)
ref_l, ref_n = S.ccl_label(batch, planes)
bad_lines = set()
nbad = 0
for r in range(reps):
    l, k = S.ccl_label(batch, planes)
    if not bool((l == ref_l).all()):
        nbad += 1
        if nbad > 3:
            continue
        for i in range(n):
            a, b = batch.plane(l, i), batch.plane(ref_l, i)
            if not bool((a == b).all()):
                bad_lines.add(i)
                d = (a != b).nonzero()
                print(f"rep {r}: line {i} (W={widths[i]}) differs at {d.shape[0]} px, first {d[0].tolist()}, cols {int(d[:,1].min())}..{int(d[:,1].max())}, "
                      f"num {int(k[i])} vs {int(ref_n[i])}")
print("lines differing between runs:", sorted(bad_lines))
wrong = []
for i in range(n):
    nn, ref = cv2.connectedComponents(masks[i])
    if int(ref_n[i]) != nn or not np.array_equal(batch.plane(ref_l, i).cpu().numpy(), ref):
        wrong.append(i)
print("lines differing from cv2 (first run):", wrong)

print("runs differing from the first:", nbad, "of", reps)
