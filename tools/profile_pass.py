"""One warmed pass of the hot path for ncu: UNet over `--tiles` tiles, then the bandwidth-bound
stages (tile_extract, glue, CCL, stats, group canvases) over `--lines` synthetic lines.

Only the region between cudaProfilerStart/Stop is meant to be captured:
    ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/x \
        python tools/profile_pass.py --what unet
Without ncu it prints per-stage CUDA-event times (the numbers to quote; never quote times taken under ncu).
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from stroke_derenderer_b200 import _lib, segment as S  # noqa: E402
from stroke_derenderer_b200.engine import UNetEngine  # noqa: E402
from stroke_derenderer_b200.synth import config_widths, synth_dense_mask, synth_line  # noqa: E402
from stroke_derenderer_b200.weights import make_parity_weights  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="all", choices=["all", "unet", "seg", "dense"])
    ap.add_argument("--tiles", type=int, default=128)
    ap.add_argument("--lines", type=int, default=64)
    ap.add_argument("--sustain-reps", type=int, default=60, help="back-to-back UNet passes of the power-capped steady-state timing")
    args = ap.parse_args()
    _lib.require_cuda()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    gold = json.loads((ROOT / "tests" / "golden" / "golden.json").read_text())
    state = make_parity_weights(gold["unet"]["weights_seed"])
    state["Conv_1x1.bias"] = np.array([gold["unet"]["head_bias"]], np.float32)

    widths = config_widths(512)[:args.lines]
    images = [synth_line(int(w), seed=i) for i, w in enumerate(widths)]
    batch = S.plan_batch([im.shape[1] for im in images], dev)
    d_rgb = S.pack_lines_rgb(images, batch).to(dev)
    tiles = S.tile_extract_f16(batch, d_rgb)
    nt = min(args.tiles, batch.n_tiles)
    out = {}

    def ev(fn, reps=3):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    if args.what in ("all", "unet"):
        eng = UNetEngine(state, device=0, max_tiles=nt)
        masks = torch.empty((nt, 128, 384), dtype=torch.uint8, device=dev)
        if batch.n_tiles < nt:
            raise SystemExit(f"--lines {args.lines} give {batch.n_tiles} tiles, fewer than --tiles {nt}")
        out["unet_ms"] = ev(lambda: eng.forward_into(tiles[:nt], masks, 0.5))
        out["unet_sustained_ms"] = ev(lambda: eng.forward_into(tiles[:nt], masks, 0.5), reps=args.sustain_reps)   # power-capped steady state
        for small in (37, 64):
            if small < nt:
                out[f"unet_ms_{small}tiles"] = ev(lambda: eng.forward_into(tiles[:small], masks, 0.5), reps=5)
        torch.cuda.profiler.start()
        eng.forward_into(tiles[:nt], masks, 0.5)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        eng.enable_timing(True)
        eng.forward_into(tiles[:nt], masks, 0.5)
        out["unet_layers_ms"] = {n: round(t, 4) for n, t in eng.layer_times()}
        eng.enable_timing(False)
        eng.close()

    if args.what in ("all", "seg", "dense"):
        if args.what == "dense":
            # BASELINE config 5: 128x16384 masks with thousands of islands per line
            n = max(args.lines, 16)
            batch = S.plan_batch([16384] * n, dev)
            planes_h = np.zeros(batch.px_total, np.uint8)
            for i in range(n):
                ln = batch.lines[i]
                m = synth_dense_mask(16384, 0.003 if i % 2 == 0 else 0.01, seed=i) * 255
                planes_h[int(ln["px_off"]):int(ln["px_off"]) + 128 * int(ln["pitch"])] = m.reshape(-1)
            planes = torch.from_numpy(planes_h).to(dev)
        else:
            # text-like masks: the ink of the synthetic lines (there is no glue pass any more: the UNet head writes the planes)
            planes_h = np.zeros(batch.px_total, np.uint8)
            for im, ln in zip(images, batch.lines):
                off, pitch = int(ln["px_off"]), int(ln["pitch"])
                planes_h[off:off + 128 * pitch].reshape(128, pitch)[:, :im.shape[1]] = (im[:, :, 0] < 128) * 255
            planes = torch.from_numpy(planes_h).to(dev)
            out["tile_extract_f16_ms"] = ev(lambda: S.tile_extract_f16(batch, d_rgb, out=tiles))
        seg = S.Segmenter(None, device=dev)
        work = torch.empty(_lib.lib().sd_ccl_workspace_bytes(batch.blk_total, batch.n_lines), dtype=torch.uint8, device=dev)
        out["ccl_ms"] = ev(lambda: S.ccl_label(batch, planes, work))
        out["ccl_label_stats_ms"] = ev(lambda: S.ccl_label_stats(batch, planes, max(S.stats_capacity(batch), 2_000_000 if args.what == "dense" else 0), work))
        seg.partition(batch, planes, canvases="device")
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        if args.what != "dense":
            S.tile_extract_f16(batch, d_rgb, out=tiles)
        res = seg.partition(batch, planes, canvases="device", crops=True)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        if len(res["groups"]):
            out["group_crops_ms"] = ev(lambda: S.group_crops(dev, res["canvas"], res["_keep"][0], res["groups"]))
            out["group_crops_f32_ms"] = ev(lambda: S.group_crops(dev, res["canvas"], res["_keep"][0], res["groups"],
                                                                 lut=S.input_lut([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])))
        px = 128 * int(sum(batch.widths))
        out["seg"] = {"lines": batch.n_lines, "px": px, "islands": int(res["num"].sum() - batch.n_lines),
                      "groups": int(len(res["groups"])), "ccl_GBps_algorithmic": 5 * px / out["ccl_ms"] / 1e6}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
