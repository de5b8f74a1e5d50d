"""Where the end-to-end step loses time against the resident step: the same 512-line job (BASELINE config 3) timed as
resident_step, host_step from pre-packed pinned lines, host_step packing numpy lines inside, and host_step into the
gather arena.  CUDA events + synchronize around K steps each, after warm-up.  One JSON line."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from stroke_derenderer_b200 import gather as G  # noqa: E402
from stroke_derenderer_b200.engine import UNetEngine  # noqa: E402
from stroke_derenderer_b200.pipeline import LineSegmentationJob  # noqa: E402
from stroke_derenderer_b200.synth import config_widths, synth_line  # noqa: E402
from stroke_derenderer_b200.weights import make_parity_weights  # noqa: E402


def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    gold = json.loads((ROOT / "tests" / "golden" / "golden.json").read_text())
    state = make_parity_weights(gold["unet"]["weights_seed"])
    state["Conv_1x1.bias"] = np.array([gold["unet"]["head_bias"]], np.float32)
    widths = config_widths(512)
    images = [synth_line(int(w), seed=i) for i, w in enumerate(widths)]
    eng = UNetEngine(state, device=0, max_tiles=256)
    pre = LineSegmentationJob(eng, images, prepack=True)
    raw = LineSegmentationJob(eng, images, prepack=False)
    arena = G.ResultArena(f"sd_e2e_variants_{os.getpid()}", [raw.arena_bytes()], 0, create=True)
    arena.register()
    wr = G.RegionWriter(arena.region(0), len(raw.chunks))

    def timed(fn):
        keep = None
        for _ in range(3):
            keep = fn()
        keep = None
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(K):
            keep = fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / K

    out = {}
    for rnd in range(2):            # twice, in both orders: thermal / power state moves the numbers by a percent or two
        order = [("resident", pre.resident_step), ("host_prepacked", pre.host_step), ("host_pack_inside", raw.host_step),
                 ("host_pack_inside_arena", lambda: raw.host_step(wr))]
        if rnd:
            order.reverse()
        for name, fn in order:
            out.setdefault(name, []).append(round(timed(fn), 1))
    print(json.dumps(out))
    del wr
    arena.close()
    eng.close()


if __name__ == "__main__":
    main()
