"""Whole-job driver: shards line images across the GPUs of one box (one process per
GPU, no data-path collective: tiles and lines are independent) and
runs the batched segmentation step on each.

The hot-path step of one rank:
   [H2D lines] -> tile_extract -> Attention-UNet (tcgen05) -> glue -> CCL -> stats
   -> [D2H counts+stats] -> host interval grouping -> group canvases -> 224x224 crops -> [D2H results]
"""

from __future__ import annotations

import os
import sys
import time

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


class _Chunk:
    pass


class LineSegmentationJob:
    """One rank's share of a job, cut into chunks of lines that flow through three streams:
    copy (H2D of packed lines) -> unet (tile_extract, Attention-UNet, glue) -> part (CCL, stats,
    host grouping, canvases, D2H).  The UNet stream never waits for the host: while the host
    clusters the islands of chunk k, the tensor cores are already on chunk k+1.

    `resident_step` times the hot path with inputs already in HBM; `host_step` is the call a
    user makes with (pinned) host buffers: H2D of the lines and D2H of masks, stats and group
    canvases are inside it."""

    def __init__(self, engine: UNetEngine, images, bin_thr: float = 0.5, lines_per_chunk: int = 64, crops: bool = True):
        self.engine = engine
        self.crops = crops                  # also build the 224x224 stroke-estimator crops of every group
        self.device = engine.device
        self.seg = S.Segmenter(engine, bin_thr=bin_thr)
        self.chunks = []
        with torch.cuda.device(self.device):
            self.s_copy, self.s_unet, self.s_part = (torch.cuda.Stream(self.device) for _ in range(3))
            t0 = 0
            for c0 in range(0, len(images), lines_per_chunk):
                imgs = images[c0:c0 + lines_per_chunk]
                ch = _Chunk()
                ch.index = len(self.chunks)
                ch.batch = S.plan_batch([S.resized_width(im) for im in imgs], self.device)
                ch.h_rgb = S.pack_lines_rgb(imgs, ch.batch, pinned=True)
                ch.d_rgb = ch.h_rgb.to(self.device)
                ch.d_rgb_in = torch.empty_like(ch.d_rgb)
                ch.resize = S.ResizePlan(imgs, ch.batch)            # lines whose height is not 128 (none in the configs)
                ch.resize.upload()
                ch.resize.run(ch.d_rgb)
                ch.t0, ch.t1 = t0, t0 + ch.batch.n_tiles        # the chunk's range in the job-wide tile stack
                t0 = ch.t1
                ch.planes = torch.empty(ch.batch.px_total, dtype=torch.uint8, device=self.device)
                ch.h_planes = torch.empty(ch.batch.px_total, dtype=torch.uint8, pin_memory=True)
                self.chunks.append(ch)
            # one tile stack for the whole job: UNet batches run across chunk boundaries, so only the
            # last batch of the job is partial
            self.tiles = torch.empty((max(t0, 1), TILE_H, TILE_W, 8), dtype=torch.float16, device=self.device)
            # job-wide sd_tile_dst table: tile k of the stack pastes into its chunk's planes (glue fused into the head)
            self.dst = torch.cat([ch.batch.tile_dst(ch.planes) for ch in self.chunks]) if self.chunks else None
            for ch in self.chunks:
                ch.tiles = self.tiles[ch.t0:ch.t1]
        self.n_tiles = sum(c.batch.n_tiles for c in self.chunks)
        self.n_lines = sum(c.batch.n_lines for c in self.chunks)

    def _run(self, from_host: bool, canvases: str):
        with torch.cuda.device(self.device):
            cur = torch.cuda.current_stream(self.device)
            for st in (self.s_copy, self.s_unet, self.s_part):
                st.wait_stream(cur)
            ready = [None] * len(self.chunks)
            mt = self.engine.max_tiles
            done = 0                      # tiles binarized so far
            glued = 0                     # chunks glued so far
            for k, ch in enumerate(self.chunks):
                src = ch.d_rgb
                if from_host:
                    with torch.cuda.stream(self.s_copy):
                        ch.d_rgb_in.copy_(ch.h_rgb, non_blocking=True)
                        ch.resize.upload()
                        ev = torch.cuda.Event(); ev.record(self.s_copy)
                    self.s_unet.wait_event(ev)
                    src = ch.d_rgb_in
                with torch.cuda.stream(self.s_unet):
                    if from_host:
                        ch.resize.run(src)
                    S.tile_extract_f16(ch.batch, src, out=ch.tiles)
                    ch.planes.zero_()                 # the head ORs foreground bytes into the planes
                    last = k == len(self.chunks) - 1
                    while done + mt <= ch.t1 or (last and done < ch.t1):
                        n = min(mt, self.n_tiles - done)
                        self.engine.forward_lines(self.tiles[done:done + n], self.dst[16 * done:16 * (done + n)], self.seg.bin_thr)
                        done += n
                    ev = None
                    while glued < len(self.chunks) and self.chunks[glued].t1 <= done:   # chunks whose last tile is through
                        if ev is None:
                            ev = torch.cuda.Event(); ev.record(self.s_unet)
                        ready[glued] = ev
                        glued += 1
            results = []
            dbg = os.environ.get("SD_PIPE_DEBUG")
            t_dbg = [time.perf_counter()]
            with torch.cuda.stream(self.s_part):
                for ch, ev in zip(self.chunks, ready):
                    self.s_part.wait_event(ev)
                    if from_host:
                        ch.h_planes.copy_(ch.planes, non_blocking=True)
                    res = self.seg.partition(ch.batch, ch.planes, canvases=canvases, key=("chunk", ch.index),
                                             zero_copy=True, crops=self.crops)
                    if from_host and self.crops and res["crops"] is not None:
                        img = res["crops"]["image"]
                        hb = S.pinned_buffer((("chunk", ch.index), "crops"), img.numel())[:img.numel()]
                        hb.copy_(img.view(-1), non_blocking=True)
                        res["crops"]["image_host"] = hb.numpy().reshape(tuple(img.shape))
                        res["crops"]["input_host"] = None
                    results.append(res)
                    t_dbg.append(time.perf_counter())
            cur.wait_stream(self.s_unet)
            cur.wait_stream(self.s_part)
            if from_host:
                cur.synchronize()
            if dbg:
                t_dbg.append(time.perf_counter())
                print("[pipe] enqueue->partitions done (ms):", [round(1e3 * (b - a), 1) for a, b in zip(t_dbg, t_dbg[1:])],
                      file=sys.stderr, flush=True)
        return results

    def resident_step(self):
        return self._run(False, "device")

    def host_step(self):
        return self._run(True, "host")

    def h2d_bytes(self):
        return int(sum(c.batch.plan.img_bytes + (c.resize.h_src.numel() if c.resize.n else 0) for c in self.chunks))

    def d2h_bytes(self, results):
        n = sum(c.batch.px_total for c in self.chunks)
        n += sum(r["num"].nbytes + r["stats"].nbytes + r["canvas_bytes"] for r in results)
        n += sum(r["crops"]["image"].numel() for r in results if r.get("crops") is not None)
        return int(n)
