"""Where the time of the reference-signature calls goes (GPU box): get_partitions_batch over chunk sizes and host lanes,
with a per-phase breakdown of one lane (SD_PARTS_DEBUG=1).  python tools/api_profile.py [--lines 512]"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from stroke_derenderer_b200.evaluate_strokes import StrokeEstimationSession   # noqa: E402
from stroke_derenderer_b200.synth import ink_mask, synth_line                 # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--lines", type=int, default=512)
a = ap.parse_args()
rng = np.random.default_rng(1)
widths = rng.integers(1536, 6145, a.lines)
bins = [ink_mask(synth_line(int(w), seed=i)) > 0 for i, w in enumerate(widths)]
se = StrokeEstimationSession(device=0)
out = {}
for lpc in (32, 64, 128):
    for lanes in (1, 2, 3):
        se.get_partitions_batch(bins, lines_per_chunk=lpc, lanes=lanes)
        ts = []
        keep = None
        for _ in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            keep = se.get_partitions_batch(bins, lines_per_chunk=lpc, lanes=lanes)
            torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
        out[f"chunk{lpc}_lanes{lanes}"] = [round(t, 1) for t in ts]
        print(lpc, lanes, out[f"chunk{lpc}_lanes{lanes}"], sum(len(p) for p in keep), file=sys.stderr, flush=True)
print(json.dumps(out))
