"""CCL timing / profiling target: the 128-line text sample of bench.py's hbm_stages and 64 dense config-5 lines.
Without a profiler: CUDA-event times of sd_ccl_label and sd_ccl_label_stats (round 1's kernels, removed since, took 0.227 / 0.259 ms on the same two samples).  Under ncu only the region between
cudaProfilerStart/Stop is captured (one call per workload):
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv python tools/ccl_bench.py
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from stroke_derenderer_b200 import _lib, segment as S  # noqa: E402
from stroke_derenderer_b200.synth import config_widths, ink_mask, synth_dense_mask, synth_line  # noqa: E402


def pack(masks, dev):
    batch = S.plan_batch([m.shape[1] for m in masks], dev)
    host = np.zeros(batch.px_total, np.uint8)
    for m, ln in zip(masks, batch.lines):
        off, pitch = int(ln["px_off"]), int(ln["pitch"])
        host[off:off + 128 * pitch].reshape(128, pitch)[:, :m.shape[1]] = m * 255
    return batch, torch.from_numpy(host).to(dev)


def ev(fn, reps=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    import os
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    stages = "--stages" in sys.argv
    n_text = int(args[0]) if args else 128
    widths = config_widths(512)[:n_text]
    text = [ink_mask(synth_line(int(w), seed=i)) for i, w in enumerate(widths)]
    dense = [synth_dense_mask(16384, 0.003 if i % 2 == 0 else 0.01, seed=i) for i in range(64)]
    out = {}
    for name, masks in (("text", text), ("dense", dense)):
        name0 = name
        batch, planes = pack(masks, dev)
        work = torch.empty(_lib.lib().sd_ccl_workspace_bytes(batch.blk_total, batch.n_lines), dtype=torch.uint8, device=dev)
        cap = max(S.stats_capacity(batch), 2_000_000 if name0 == "dense" else 0)
        px = 128 * int(sum(m.shape[1] for m in masks))
        t = ev(lambda: S.ccl_label(batch, planes, work))
        ts = ev(lambda: S.ccl_label_stats(batch, planes, cap, work))
        if stages:                 # cumulative CUDA-event time of the first k kernels of sd_ccl_label (SD_CCL_STAGE)
            cum = []
            for k in range(1, 5):
                os.environ["SD_CCL_STAGE"] = str(k)
                cum.append(ev(lambda: S.ccl_label(batch, planes, work)))
            os.environ.pop("SD_CCL_STAGE")
            names = ["label", "seam_merge", "line (mark + scan)", "write"]
            out[name + "_stages_us"] = {n: round(1e3 * (c - p), 1) for n, c, p in zip(names, cum, [0.0] + cum[:-1])}
        out[name] = {"px": px, "label_ms": t, "label_frac": 5 * px / t / 1e6 / 6451.5, "label_stats_ms": ts,
                     "label_stats_frac": 5 * px / ts / 1e6 / 6451.5}
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        S.ccl_label_stats(batch, planes, cap, work)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
