"""Debug: one 128x128 strip of line 127 (cols 2560..2687), labelled many times (many copies per launch)."""
import sys
from pathlib import Path
import numpy as np, torch, cv2
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
from stroke_derenderer_b200 import segment as S
from stroke_derenderer_b200.synth import config_widths, ink_mask, synth_line
w = config_widths(512)
full = ink_mask(synth_line(int(w[127]), seed=127))
x0 = int(sys.argv[1]) if len(sys.argv) > 1 else 2560
wd = int(sys.argv[2]) if len(sys.argv) > 2 else 128
copies = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
m = np.ascontiguousarray(full[:, x0:x0 + wd])
n_ref, ref = cv2.connectedComponents(m)
dev = torch.device("cuda", 0)
batch = S.plan_batch([wd] * copies, dev)
host = np.zeros(batch.px_total, np.uint8)
for ln in batch.lines:
    off, pitch = int(ln["px_off"]), int(ln["pitch"])
    host[off:off + 128 * pitch].reshape(128, pitch)[:, :wd] = m
planes = torch.from_numpy(host).to(dev)
bad = 0
for r in range(5):
    l, k = S.ccl_label(batch, planes)
    kk = k.cpu().numpy()
    bad += int((kk != n_ref).sum())
print(f"strip x0={x0} w={wd}: cv2 components {n_ref - 1}; wrong copies {bad} of {5 * copies}")
