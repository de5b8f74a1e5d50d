"""Per-line mask agreement of the fused device path against the torch-CPU fp32 oracle over the first N lines of
BASELINE config 3 (default 64 lines, ~800 tiles, a couple of minutes of CPU): the distribution and the PER-LINE
MINIMUM, which is what the north star's 99.9 % bar is held against.  Writes one JSON line.
    python tools/parity_sweep.py [n_lines] > profiles/r02_parity_sweep.json
The oracle is test infrastructure (oracle/): this tool is a checker, like tests/.
"""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import segmentation_ref as O  # noqa: E402
from stroke_derenderer_b200 import segment as S  # noqa: E402
from stroke_derenderer_b200.engine import UNetEngine  # noqa: E402
from stroke_derenderer_b200.synth import config_widths, synth_line  # noqa: E402
from stroke_derenderer_b200.weights import make_parity_weights  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    gold = json.loads((ROOT / "tests" / "golden" / "golden.json").read_text())
    state = make_parity_weights(gold["unet"]["weights_seed"])
    state["Conv_1x1.bias"] = np.array([gold["unet"]["head_bias"]], np.float32)
    widths = config_widths(512)[:n]
    lines = [synth_line(int(w), seed=i) for i, w in enumerate(widths)]
    eng = UNetEngine(state, device=0, max_tiles=256)
    batch, planes = S.Segmenter(eng).binarize(lines)
    torch.cuda.synchronize()
    ort = O.TorchOrtSession(state)
    bs = O.BinarizationSessionRef()
    agree, fg, t0 = [], [], time.time()
    for i, line in enumerate(lines):
        mask = batch.plane(planes, i).cpu().numpy()
        ref = bs.binarize_image(line, ort)[:, :, 0]
        agree.append(float(((mask > 127) == (ref > 127)).mean()))
        fg.append(float((ref > 127).mean()))
    eng.close()
    a = np.array(agree)
    print(json.dumps({"lines": n, "tiles": int(batch.n_tiles), "widths": [int(w) for w in widths], "bar": 0.999,
                      "per_line_min": float(a.min()), "per_line_max": float(a.max()), "mean": float(a.mean()),
                      "lines_below_bar": int((a < 0.999).sum()), "percentiles": {str(p): float(np.percentile(a, p)) for p in (1, 5, 25, 50, 75)},
                      "fg_ref_mean": float(np.mean(fg)), "per_line": [round(x, 6) for x in agree],
                      "oracle": "torch-CPU fp32 of the published topology (unpinned)", "oracle_seconds": round(time.time() - t0, 1)}))


if __name__ == "__main__":
    main()
