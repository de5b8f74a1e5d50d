"""Debug: find the component that an erroneous CCL run splits, and where."""
import sys
from pathlib import Path
import numpy as np, torch, cv2
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
from stroke_derenderer_b200 import segment as S
from stroke_derenderer_b200.synth import config_widths, ink_mask, synth_line
n, reps = 160, 200
widths = config_widths(512)[:n]
masks = [ink_mask(synth_line(int(w), seed=i)) for i, w in enumerate(widths)]
dev = torch.device("cuda", 0)
batch = S.plan_batch([m.shape[1] for m in masks], dev)
host = np.zeros(batch.px_total, np.uint8)
for m, ln in zip(masks, batch.lines):
    off, pitch = int(ln["px_off"]), int(ln["pitch"])
    host[off:off + 128 * pitch].reshape(128, pitch)[:, :m.shape[1]] = m
planes = torch.from_numpy(host).to(dev)
refs = {i: cv2.connectedComponents(masks[i])[1] for i in (127, 156, 33)}
seen = 0
for r in range(reps):
    l, k = S.ccl_label(batch, planes)
    for i, ref in refs.items():
        got = batch.plane(l, i).cpu().numpy()
        if np.array_equal(got, ref):
            continue
        seen += 1
        for lab in np.unique(ref[ref > 0]):
            w = np.unique(got[ref == lab])
            if len(w) > 1:
                ys, xs = np.nonzero(ref == lab)
                print(f"rep {r} line {i}: cv2 component {lab} bbox x[{xs.min()},{xs.max()}] y[{ys.min()},{ys.max()}] split into GPU labels {w.tolist()}")
                for ww in w:
                    yy, xx = np.nonzero((got == ww) & (ref == lab))
                    print(f"    part {ww}: x[{xx.min()},{xx.max()}] y[{yy.min()},{yy.max()}] {len(xx)} px")
                break
        if seen >= 6:
            sys.exit(0)
print("erroneous (line, run) pairs seen:", seen)
