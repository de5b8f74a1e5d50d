"""Whole-job driver: shards line images across the GPUs of one box (one process per
GPU, no data-path collective: tiles and lines are independent) and
runs the batched segmentation step on each.

The hot-path step of one rank:
   [pack + H2D lines] -> tile_extract -> Attention-UNet (tcgen05; glue + thresholds fused into its head) -> CCL
   -> stats -> [D2H counts+stats] -> host interval grouping -> group canvases -> 224x224 crops -> [D2H results]
Results are gathered on the host: every rank's D2H copies land in its region of one shared-memory arena
(`gather.ResultArena`), the caller reads all lines in input order (reference: one process, one list of images in,
results in input order out, /root/reference/main.py:91-136).
"""

from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

from . import gather as G
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
    """Object gather of small per-line results into input order (rank 0 gets the list).  Bulk results (masks,
    crops) go through `gather.ResultArena` instead."""
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
    copy (H2D of packed lines) -> unet (tile_extract, Attention-UNet with the glue in its head) -> part (CCL,
    stats, host grouping, canvases, crops, D2H).  The UNet stream never waits for the host: while the host
    clusters the islands of chunk k, the tensor cores are already on chunk k+1.

    `resident_step` times the hot path with inputs already in HBM; `host_step` is the call a user makes with
    host buffers: packing (when `prepack=False`), H2D of the lines and D2H of masks, stats, group tables,
    canvases and crops are inside it.  With a `gather.RegionWriter` the host copies land in the rank's region of
    the shared gather arena."""

    def __init__(self, engine: UNetEngine, images, bin_thr: float = 0.5, lines_per_chunk: int = 32, crops: bool = True,
                 prepack: bool = True, seg: S.Segmenter | None = None):
        self.engine = engine
        self.crops = crops                  # also build the 224x224 stroke-estimator crops of every group
        self.device = engine.device
        self.seg = seg if seg is not None else S.Segmenter(engine, bin_thr=bin_thr)
        self.staging = self.seg.staging
        self.prepack = prepack
        self.lines_per_chunk = lines_per_chunk
        self.chunks = []
        with torch.cuda.device(self.device):
            self.s_copy, self.s_unet, self.s_part = self.seg.streams()
            t0 = 0
            for c0 in range(0, len(images), lines_per_chunk):
                imgs = images[c0:c0 + lines_per_chunk]
                ch = _Chunk()
                ch.index = len(self.chunks)
                ch.imgs = imgs
                ch.batch = S.plan_batch([S.resized_width(im) for im in imgs], self.device)
                ch.h_rgb = self.staging.get_tensor((("chunk", ch.index), "rgb"), int(ch.batch.plan.img_bytes))
                ch.d_rgb_in = torch.empty(int(ch.batch.plan.img_bytes), dtype=torch.uint8, device=self.device)
                ch.resize = S.ResizePlan(imgs, ch.batch)            # lines whose height is not 128 (none in the configs)
                ch.d_rgb = None
                if prepack:
                    S.pack_lines_rgb(imgs, ch.batch, out=ch.h_rgb)
                    ch.d_rgb = ch.h_rgb.to(self.device)
                    ch.resize.upload()
                    ch.resize.run(ch.d_rgb)
                ch.t0, ch.t1 = t0, t0 + ch.batch.n_tiles        # the chunk's range in the job-wide tile stack
                t0 = ch.t1
                ch.planes = torch.empty(ch.batch.px_total, dtype=torch.uint8, device=self.device)
                self.chunks.append(ch)
            # one tile stack for the whole job: UNet batches run across chunk boundaries, so only the
            # last batch of the job is partial
            self.tiles = torch.empty((max(t0, 1), TILE_H, TILE_W, 8), dtype=S._lib.torch_dtype(), device=self.device)
            # job-wide sd_tile_dst table: tile k of the stack pastes into its chunk's planes (glue fused into the head)
            self.dst = torch.cat([ch.batch.tile_dst(ch.planes) for ch in self.chunks]) if self.chunks else None
            for ch in self.chunks:
                ch.tiles = self.tiles[ch.t0:ch.t1]
        self.n_tiles = sum(c.batch.n_tiles for c in self.chunks)
        self.n_lines = sum(c.batch.n_lines for c in self.chunks)
        self.px_total = sum(c.batch.px_total for c in self.chunks)

    def _run(self, from_host: bool, canvases: str, writer: G.RegionWriter | None = None, collect=None, fresh: bool = False):
        with torch.cuda.device(self.device):
            cur = torch.cuda.current_stream(self.device)
            for st in (self.s_copy, self.s_unet, self.s_part):
                st.wait_stream(cur)
            if writer is not None:
                writer.begin_step()
            ready = [None] * len(self.chunks)
            mt = self.engine.max_tiles
            done = 0                      # tiles binarized so far
            glued = 0                     # chunks whose planes are complete
            for k, ch in enumerate(self.chunks):
                src = ch.d_rgb
                if from_host:
                    if not self.prepack:              # the caller's numpy images -> page-locked staging (host memcpy)
                        S.pack_lines_rgb(ch.imgs, ch.batch, out=ch.h_rgb)
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
            # fresh: results land in new page-locked arrays the caller keeps (no copy out of a reused staging buffer)
            st = writer if writer is not None else (S.FreshPinned() if fresh else self.staging)
            with torch.cuda.stream(self.s_part):
                for ch, ev in zip(self.chunks, ready):
                    self.s_part.wait_event(ev)
                    key = ("chunk", ch.index)
                    h_planes = None
                    if from_host:
                        h_planes = S.copy_d2h(st.get((key, "planes"), ch.batch.px_total), ch.planes, self.device)
                    res = self.seg.partition(ch.batch, ch.planes, canvases=canvases, key=key, zero_copy=True,
                                             crops=self.crops, staging=st, crops_to_host=from_host)
                    res["planes_host"] = h_planes
                    if collect is not None:           # per-line outputs of this chunk, built while the GPU is on later chunks
                        if canvases != "host":
                            self.s_part.synchronize()
                        collect(ch, res)
                    if writer is not None:
                        writer.put(ch.index, "groups", res["groups"].reshape(-1, 6))
                        writer.put(ch.index, "lgs", res["line_group_start"])
                        writer.set(ch.index, n_lines=ch.batch.n_lines, px_total=ch.batch.px_total, n_rows=len(res["stats"]),
                                   n_groups=len(res["groups"]), canvas_bytes=res["canvas_bytes"], crop_size=S.IMG_SIZE)
                    results.append(res)
                    t_dbg.append(time.perf_counter())
            cur.wait_stream(self.s_unet)
            cur.wait_stream(self.s_part)
            if from_host:
                cur.synchronize()
                if writer is not None:            # every byte of the step has landed in the arena
                    for ch in self.chunks:
                        writer.set(ch.index, done=1)
                    writer.end_step()
            if dbg:
                t_dbg.append(time.perf_counter())
                print("[pipe] enqueue->partitions done (ms):", [round(1e3 * (b - a), 1) for a, b in zip(t_dbg, t_dbg[1:])],
                      file=sys.stderr, flush=True)
        return results

    def binarize_step(self):
        """Only the binarization half, from the caller's numpy images: -> [ (128, W', 1) u8 {0,255} ] in line order
        (evaluate_binarize.py:130-140).  One packed D2H per chunk, straight into a fresh page-locked array
        (`segment.FreshPinned`): a line's mask is a view into it, no second host copy."""
        out = []
        fresh = S.FreshPinned()
        with torch.cuda.device(self.device):
            cur = torch.cuda.current_stream(self.device)
            for st in (self.s_copy, self.s_unet, self.s_part):
                st.wait_stream(cur)
            mt = self.engine.max_tiles
            done = glued = 0
            evs = [None] * len(self.chunks)
            hosts = [None] * len(self.chunks)
            for k, ch in enumerate(self.chunks):
                if not self.prepack:
                    S.pack_lines_rgb(ch.imgs, ch.batch, out=ch.h_rgb)
                with torch.cuda.stream(self.s_copy):
                    ch.d_rgb_in.copy_(ch.h_rgb, non_blocking=True)
                    ch.resize.upload()
                    ev = torch.cuda.Event(); ev.record(self.s_copy)
                self.s_unet.wait_event(ev)
                with torch.cuda.stream(self.s_unet):
                    ch.resize.run(ch.d_rgb_in)
                    S.tile_extract_f16(ch.batch, ch.d_rgb_in, out=ch.tiles)
                    ch.planes.zero_()
                    last = k == len(self.chunks) - 1
                    while done + mt <= ch.t1 or (last and done < ch.t1):
                        n = min(mt, self.n_tiles - done)
                        self.engine.forward_lines(self.tiles[done:done + n], self.dst[16 * done:16 * (done + n)], self.seg.bin_thr)
                        done += n
                    while glued < len(self.chunks) and self.chunks[glued].t1 <= done:
                        g = self.chunks[glued]
                        hosts[glued] = S.copy_d2h(fresh.get(None, g.batch.px_total), g.planes, self.device)
                        ev = torch.cuda.Event(); ev.record(self.s_unet)
                        evs[glued] = ev
                        glued += 1
            for ch, ev, hp in zip(self.chunks, evs, hosts):
                ev.synchronize()
                for ln in ch.batch.lines:
                    off, pitch, w = int(ln["px_off"]), int(ln["pitch"]), int(ln["width"])
                    out.append(hp[off:off + TILE_H * pitch].reshape(TILE_H, pitch)[:, :w, None])
            cur.wait_stream(self.s_unet)
        return out

    def resident_step(self):
        if not self.prepack:
            raise RuntimeError("resident_step needs a job built with prepack=True")
        return self._run(False, "device")

    def host_step(self, writer: G.RegionWriter | None = None):
        return self._run(True, "host", writer)

    def h2d_bytes(self):
        return int(sum(c.batch.plan.img_bytes + (c.resize.h_src.numel() if c.resize.n else 0) for c in self.chunks))

    def d2h_bytes(self, results):
        n = sum(c.batch.px_total for c in self.chunks)
        n += sum(r["num"].nbytes + r["stats"].nbytes + r["canvas_bytes"] for r in results)
        n += sum(r["crops"]["image"].numel() for r in results if r.get("crops") is not None)
        return int(n)

    def arena_bytes(self) -> int:
        """Capacity of this rank's gather-arena region for one step."""
        return G.region_capacity(self.n_tiles, self.n_lines, self.px_total, max(len(self.chunks), 1))

    # ---- per-line outputs in the reference's shapes ------------------------------------------------------------
    def chunk_outputs(self, ch, res, lut, copy: bool = True):
        """([mask (128, W', 1) u8 {0,255}], [[partition dict]]) for the lines of one finished chunk — what
        `BinarizationSession.binarize_images` (evaluate_binarize.py:130-140) and
        `StrokeEstimationSession.get_partitions` (evaluate_strokes.py:186-224) return.  `image_input` of a
        partition is materialised on first access (`segment.LazyPartition`)."""
        hp = res["planes_host"]
        cr = res["crops"]
        img_host = cr["image_host"] if cr is not None and "image_host" in cr else None
        if copy:                          # fresh arrays (two host threads; numpy copies release the GIL)
            f1 = S.host_pool().submit(np.copy, hp)
            if img_host is not None:
                img_host = img_host.copy()
            hp = f1.result()
        masks = []
        for ln in ch.batch.lines:
            off, pitch, w = int(ln["px_off"]), int(ln["pitch"]), int(ln["width"])
            masks.append(hp[off:off + TILE_H * pitch].reshape(TILE_H, pitch)[:, :w, None])
        parts = S.build_partitions(lut, img_host, res["groups"], cr, res["line_group_start"], ch.batch.n_lines)
        return masks, parts

    def line_outputs(self, results, copy: bool = True, mean=None, std=None):
        """After `host_step`: per-line masks and partitions of the whole job, in this job's line order."""
        lut = S.input_lut(mean if mean is not None else S.IMAGENET_MEAN, std if std is not None else S.IMAGENET_STD)
        masks, parts = [], []
        for ch, res in zip(self.chunks, results):
            m, p = self.chunk_outputs(ch, res, lut, copy)
            masks += m; parts += p
        return masks, parts


def segment_lines(engine: UNetEngine, images, bin_thr: float = 0.5, lines_per_chunk: int = 32, seg: S.Segmenter | None = None,
                  copy: bool = True):
    """The fused public call on one GPU: a list of (h, w, 3) u8 line images (plain numpy) in, per line the
    binarized mask and the stroke-estimator partitions out, in input order — `binarize_image` + `main.py:108` +
    `get_partitions` of the reference for every image, as one pipelined job.  A mask is a (128, W', 1) view into one
    fresh page-locked array per chunk of lines, a partition's `image` a view into the chunk's crop array: the D2H copies
    land directly in what is returned (`copy=False`: in the reused staging instead, valid until the next call)."""
    t0 = time.perf_counter()
    if seg is None:
        seg = S.Segmenter.for_engine(engine, bin_thr)
    job = LineSegmentationJob(engine, images, bin_thr=bin_thr, lines_per_chunk=lines_per_chunk, crops=True, prepack=False, seg=seg)
    t1 = time.perf_counter()
    lut = S.input_lut(S.IMAGENET_MEAN, S.IMAGENET_STD)
    masks, parts = [], []
    t_collect = [0.0]

    def collect(ch, res):
        tc = time.perf_counter()
        m, p = job.chunk_outputs(ch, res, lut, copy=False)
        masks.extend(m); parts.extend(p)
        t_collect[0] += time.perf_counter() - tc
    job._run(True, "device", collect=collect, fresh=copy)
    if os.environ.get("SD_PIPE_DEBUG"):
        print(f"[segment_lines] job construction {1e3 * (t1 - t0):.1f} ms, run {1e3 * (time.perf_counter() - t1):.1f} ms "
              f"(of which per-line outputs {1e3 * t_collect[0]:.1f} ms)", file=sys.stderr, flush=True)
    return masks, parts


class ShardedSegmentation:
    """One job over all GPUs of a box, the way the reference's single process sees it (main.py:91-136: one list of
    images in, results in input order out): one PROCESS per GPU (torchrun ranks), whole lines dealt to ranks by tile
    count (`shard_lines`), no data-path collective.  Every rank runs its shard through `LineSegmentationJob.host_step`
    with its D2H copies landing in its page-locked region of one `/dev/shm` arena; after a barrier the caller (rank 0)
    holds every line of every rank in input order (`gather.GatheredResults`).

    widths: the resized widths of ALL lines (every rank knows them); images: THIS rank's lines, in the order of
    `shards[rank]`; barrier: a callable that synchronises the ranks (torch.distributed.barrier), None for one rank."""

    def __init__(self, engine: UNetEngine, images, widths, rank: int = 0, world: int = 1, barrier=None, name: str | None = None,
                 lines_per_chunk: int = 32, crops: bool = True, all_caps=None):
        self.rank, self.world, self.barrier = rank, world, barrier
        self.widths = [int(w) for w in widths]
        self.shards = shard_lines(self.widths, world)
        if len(images) != len(self.shards[rank]):
            raise ValueError(f"rank {rank}: got {len(images)} lines, its shard has {len(self.shards[rank])}")
        self.lines_per_chunk = lines_per_chunk
        self.job = LineSegmentationJob(engine, images, lines_per_chunk=lines_per_chunk, crops=crops, prepack=False)
        # every rank can size every region: the capacity depends on the shard's widths only
        self.caps = all_caps if all_caps is not None else [self._capacity(s) for s in self.shards]
        self.name = name or f"sd_b200_gather_{os.environ.get('MASTER_PORT', 'solo')}_{os.getppid() if world > 1 else os.getpid()}"
        self.arena = G.ResultArena(self.name, self.caps, rank, create=True) if rank == 0 else None
        self._sync()
        if self.arena is None:
            self.arena = G.ResultArena(self.name, self.caps, rank, create=False)
        self.arena.register()
        self.writer = G.RegionWriter(self.arena.region(rank), max(len(self.job.chunks), 1))

    @staticmethod
    def shard_of(widths, rank: int, world: int):
        """Global indices of the lines of `rank` (sorted), before the job exists."""
        return shard_lines([int(w) for w in widths], world)[rank]

    def _capacity(self, idx) -> int:
        w = [self.widths[i] for i in idx]
        tiles = sum(n_tiles_for_width(x) for x in w)
        px = sum(TILE_H * ((x + 127) // 128 * 128) for x in w)
        return G.region_capacity(tiles, len(w), px, max((len(w) + self.lines_per_chunk - 1) // self.lines_per_chunk, 1))

    def _sync(self):
        if self.world > 1 and self.barrier is not None:
            self.barrier()

    def step(self):
        """Runs this rank's shard; returns (this rank's chunk results, GatheredResults on rank 0 / None elsewhere).
        The gathered views stay valid until the next `step` (call `release()` when done reading)."""
        res = self.job.host_step(self.writer)            # ends with a stream synchronize: this rank's bytes are in the arena
        self._sync()
        got = None
        if self.rank == 0:
            got = G.GatheredResults(self.arena, self.shards, self.widths, self.lines_per_chunk, step=self.writer.step)
        return res, got

    def release(self):
        """Readers are done: the next step may reuse the arena."""
        self._sync()

    def close(self):
        self.arena.close()
