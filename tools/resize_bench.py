"""Times sd_resize_lines (general-height resize_to_height, common.py:85-93) on a batch of synthetic lines and
reports algorithmic HBM throughput: (source bytes + destination bytes) / kernel time.
  python tools/resize_bench.py [--lines 64] [--height 200] [--width 6000] [--out gpurun_out/resize_bench.json]"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from stroke_derenderer_b200 import segment as S  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--lines", type=int, default=64)
ap.add_argument("--height", type=int, default=200)
ap.add_argument("--width", type=int, default=6000)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--out", default="")
a = ap.parse_args()

dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
imgs = [rng.integers(0, 256, (a.height, a.width, 3), dtype=np.uint8) for _ in range(a.lines)]
batch = S.plan_batch([S.resized_width(im) for im in imgs], dev)
d_rgb = torch.empty(int(batch.plan.img_bytes), dtype=torch.uint8, device=dev)
rp = S.ResizePlan(imgs, batch)
rp.upload()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ms = []
for it in range(a.iters + 3):
    flush.zero_()                                     # evict L2 between launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); rp.run(d_rgb); e1.record()
    torch.cuda.synchronize()
    if it >= 3:
        ms.append(e0.elapsed_time(e1))
nbytes = sum(im.size for im in imgs) + sum(128 * int(l["width"]) * 3 for l in batch.lines)
t = float(np.median(ms))
out = {"kernel": "resize_lines_kernel", "lines": a.lines, "src": [a.height, a.width, 3], "dst_w": int(batch.lines[0]["width"]),
       "algorithmic_bytes": int(nbytes), "ms": t, "GB_per_s": nbytes / t / 1e6, "l2": "flushed between launches"}
print(json.dumps(out))
if a.out:
    Path(a.out).write_text(json.dumps(out) + "\n")
