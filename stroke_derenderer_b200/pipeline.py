"""Whole-job driver: shards line images across the GPUs of one box (one process per
GPU, no data-path collective: tiles and lines are independent) and
runs the batched segmentation step on each.

The hot-path step of one rank:
   [H2D lines] -> tile_extract -> Attention-UNet (tcgen05) -> glue -> CCL -> stats
   -> [D2H counts+stats] -> host interval grouping -> group canvases -> [D2H results]
"""

from __future__ import annotations

import numpy as np
import torch

from . import segment as S
from .engine import TILE_H, TILE_W, UNetEngine
from .synth import n_tiles_for_width


def shard_lines(widths, world: int):
    """Greedy longest-processing-time assignment of whole lines to ranks by tile
    count (a line's tiles stay on one GPU so glue/CCL need no exchange).
    Returns list (per rank) of sorted line indices; deterministic."""
    tiles = np.array([n_tiles_for_width(int(w)) for w in widths], dtype=np.int64)
    order = np.argsort(-tiles, kind="stable")
    load = np.zeros(world, np.int64)
    out = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(load))
        out[r].append(int(i))
        load[r] += tiles[i]
    return [sorted(x) for x in out]


def gather_in_order(local_results, local_indices, n_total, world, rank, group=None):
    """Host-side gather of per-line results into input order (rank 0 gets the list)."""
    import torch.distributed as dist
    if world == 1:
        out = [None] * n_total
        for i, r in zip(local_indices, local_results):
            out[i] = r
        return out
    payload = (local_indices, local_results)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(payload, gathered, dst=0, group=group)
    if rank != 0:
        return None
    out = [None] * n_total
    for idxs, ress in gathered:
        for i, r in zip(idxs, ress):
            out[i] = r
    return out


class LineSegmentationJob:
    """One rank's share of a job.  `resident_step` times the hot path with inputs already
    in HBM; `host_step` is the same call a user makes with host buffers (H2D + D2H inside)."""

    def __init__(self, engine: UNetEngine, images, bin_thr: float = 0.5):
        self.engine = engine
        self.device = engine.device
        self.seg = S.Segmenter(engine, bin_thr=bin_thr)
        self.images = images
        with torch.cuda.device(self.device):
            self.batch = S.plan_batch([im.shape[1] for im in images], self.device)
            self.h_rgb = S.pack_lines_rgb(images, self.batch, pinned=True)
            self.d_rgb = self.h_rgb.to(self.device)
            nt = self.batch.n_tiles
            self.tiles = torch.empty((nt, TILE_H, TILE_W, 8), dtype=torch.float16, device=self.device)
            self.masks = torch.empty((nt, TILE_H, TILE_W), dtype=torch.uint8, device=self.device)
            self.planes = torch.empty(self.batch.px_total, dtype=torch.uint8, device=self.device)
            self.h_planes = torch.empty(self.batch.px_total, dtype=torch.uint8, pin_memory=True)

    @property
    def n_tiles(self): return self.batch.n_tiles
    @property
    def n_lines(self): return self.batch.n_lines

    def _device_binarize(self, d_rgb):
        S.tile_extract_f16(self.batch, d_rgb, out=self.tiles)
        mt = self.engine.max_tiles
        for s in range(0, self.batch.n_tiles, mt):
            self.engine.forward_into(self.tiles[s:s + mt], self.masks[s:s + mt], self.seg.bin_thr)
        S.glue_u8(self.batch, self.masks, out=self.planes)

    def resident_step(self):
        with torch.cuda.device(self.device):
            self._device_binarize(self.d_rgb)
            return self.seg.partition(self.batch, self.planes, want_canvases=True)

    def host_step(self):
        """Host buffers in, host results out."""
        with torch.cuda.device(self.device):
            d_rgb = self.h_rgb.to(self.device, non_blocking=True)
            self._device_binarize(d_rgb)
            self.h_planes.copy_(self.planes, non_blocking=True)
            res = self.seg.partition(self.batch, self.planes, want_canvases=True)   # D2H of counts, stats, canvases inside
            torch.cuda.current_stream(self.device).synchronize()
        return res

    def h2d_bytes(self): return int(self.batch.plan.img_bytes)

    def d2h_bytes(self, res):
        n = self.batch.px_total + res["num"].nbytes + res["stats"].nbytes
        n += sum(c.size for line in res["canvases"] for c, _ in line)
        return int(n)
