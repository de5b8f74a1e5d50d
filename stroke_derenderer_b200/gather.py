"""Host-side gather of a multi-GPU segmentation job (BASELINE north star: "sharded across the
8 GPUs of one box with per-GPU streams and a host-side gather, no NCCL collective").

The reference processes one list of images in one process and returns results in input order
(/root/reference/main.py:91-136).  Here every rank (one process per GPU) owns a region of ONE shared-memory
arena (`/dev/shm/<name>`), page-locks it (sd_host_register) and lets its device-to-host copies land there
directly: masks, island counts, cv2-layout stats, group tables, group canvases and 224x224 crops.  The caller
(rank 0) maps the same file and reads every line's results in input order: the gather costs no collective and no
extra host copy, only the D2H traffic each rank produces anyway.

Region layout (all little-endian, offsets relative to the region start, blocks 256-byte aligned):
    int64 header[8]            = MAGIC, n_chunks, step id, bytes used, 0...
    int64 chunk[n_chunks][16]  = CHUNK_FIELDS below
    data blocks                (bump-allocated while the step runs)
"""

from __future__ import annotations

import mmap
import os

import numpy as np

from . import _lib

MAGIC = 0x5344423230304741           # "SDB200GA"
HDR_WORDS = 8
CH_WORDS = 16
CHUNK_FIELDS = ("n_lines", "off_planes", "px_total", "off_num", "off_stats", "n_rows", "off_groups", "n_groups",
                "off_lgs", "off_canvas", "canvas_bytes", "off_crops", "crop_size", "done")
ALIGN = 256
PAGE = 4096


def _up(v: int, a: int) -> int:
    return (int(v) + a - 1) // a * a


def region_capacity(n_tiles: int, n_lines: int, px_total: int, n_chunks: int, crop_size: int = 224,
                    groups_per_tile: float = 4.5, islands_per_tile: float = 40.0) -> int:
    """Bytes a rank's region needs for one step: exact for the planes, generous bounds for what depends on the
    data (islands, groups; measured on the synthetic lines: ~6 islands and ~2.7 groups per tile).  A step that
    outgrows its region raises, it never writes past it."""
    groups = int(groups_per_tile * n_tiles) + 64 * n_chunks
    islands = int(islands_per_tile * n_tiles) + 256 * n_chunks
    b = 8 * (HDR_WORDS + CH_WORDS * n_chunks) + px_total + 4 * n_lines + 20 * islands + 48 * groups + 8 * (n_lines + n_chunks)
    b += groups * (crop_size * crop_size + 128 * 128)         # crops + canvases (a canvas is at most 128 x 128)
    return _up(b + ALIGN * 8 * n_chunks + PAGE, PAGE)


class ResultArena:
    """One shared-memory file, one page-aligned region per rank."""

    def __init__(self, name: str, region_bytes, rank: int, create: bool):
        self.path = f"/dev/shm/{name}"
        self.rank = rank
        self.sizes = [_up(b, PAGE) for b in region_bytes]
        self.offsets = np.concatenate([[0], np.cumsum(self.sizes)]).astype(np.int64)
        total = int(self.offsets[-1])
        self.created = create
        if create:
            fd = os.open(self.path, os.O_CREAT | os.O_RDWR | os.O_TRUNC, 0o600)
            os.ftruncate(fd, total)
        else:
            fd = os.open(self.path, os.O_RDWR)
        try:
            self.mm = mmap.mmap(fd, total)
        finally:
            os.close(fd)
        self.buf = np.frombuffer(self.mm, dtype=np.uint8)
        self._registered = None

    def region(self, r: int) -> np.ndarray:
        return self.buf[int(self.offsets[r]):int(self.offsets[r]) + self.sizes[r]]

    def register(self):
        """Page-locks this rank's region so that D2H copies into it are asynchronous and run at PCIe rate."""
        if self._registered is None:
            reg = self.region(self.rank)
            _lib.check(_lib.lib().sd_host_register(reg.ctypes.data, reg.nbytes), "sd_host_register")
            self._registered = reg.ctypes.data

    def close(self):
        if self._registered is not None:
            _lib.lib().sd_host_unregister(self._registered)
            self._registered = None
        self.buf = None
        try:
            self.mm.close()
        except BufferError:            # numpy views handed out earlier are still alive: the mapping goes with them
            pass
        if self.created and os.path.exists(self.path):
            os.unlink(self.path)


class RegionWriter:
    """Bump allocator + chunk directory over one rank's region.  Implements the `staging` interface of
    `Segmenter.partition` (`get((chunk key, kind), nbytes)`): the host copies of a chunk land in the arena."""

    KINDS = {"planes": "off_planes", "num": "off_num", "stats": "off_stats", "groups": "off_groups", "lgs": "off_lgs",
             "canvas": "off_canvas", "crops": "off_crops"}

    def __init__(self, region: np.ndarray, n_chunks: int):
        self.region = region
        self.n_chunks = n_chunks
        self.hdr = region[:8 * HDR_WORDS].view(np.int64)
        self.dir = region[8 * HDR_WORDS:8 * (HDR_WORDS + CH_WORDS * n_chunks)].view(np.int64).reshape(n_chunks, CH_WORDS)
        self.data0 = _up(8 * (HDR_WORDS + CH_WORDS * n_chunks), ALIGN)
        self.top = self.data0
        self.step = 0

    def begin_step(self):
        self.step += 1
        self.top = self.data0
        self.hdr[:] = 0
        self.dir[:] = 0

    def alloc(self, nbytes: int):
        off = self.top
        end = _up(off + nbytes, ALIGN)
        if end > self.region.nbytes:
            raise MemoryError(f"gather arena region too small: need {end} bytes, have {self.region.nbytes} "
                              "(raise groups_per_tile / islands_per_tile in region_capacity)")
        self.top = end
        return off, self.region[off:off + nbytes]

    def get(self, key, nbytes: int) -> np.ndarray:
        (_, chunk), kind = key
        off, view = self.alloc(nbytes)
        self.dir[chunk, CHUNK_FIELDS.index(self.KINDS[kind])] = off
        return view

    def put(self, chunk: int, kind: str, arr: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(arr)
        view = self.get((("chunk", chunk), kind), a.nbytes)
        view[:] = a.view(np.uint8).reshape(-1)
        return view

    def set(self, chunk: int, **fields):
        for k, v in fields.items():
            self.dir[chunk, CHUNK_FIELDS.index(k)] = int(v)

    def end_step(self):
        self.hdr[0], self.hdr[1], self.hdr[2], self.hdr[3] = MAGIC, self.n_chunks, self.step, self.top


class GatheredResults:
    """What the caller (rank 0) holds after a step: every line's results, in INPUT order, as views into the arena.
    `shards[r]` = sorted global line indices of rank r (pipeline.shard_lines), `widths` = all line widths,
    lines of a rank are chunked `lines_per_chunk` at a time exactly like LineSegmentationJob does."""

    def __init__(self, arena: ResultArena, shards, widths, lines_per_chunk: int, step: int | None = None):
        self.arena = arena
        self.widths = np.asarray(widths, dtype=np.int64)
        n = len(self.widths)
        self.rank = np.full(n, -1, np.int32)
        self.chunk = np.zeros(n, np.int32)
        self.k = np.zeros(n, np.int32)
        self._dir, self._plan = [], {}
        self.shards = shards
        self.lines_per_chunk = lines_per_chunk
        for r, idx in enumerate(shards):
            reg = arena.region(r)
            hdr = reg[:8 * HDR_WORDS].view(np.int64)
            n_chunks = (len(idx) + lines_per_chunk - 1) // lines_per_chunk
            if int(hdr[0]) != MAGIC or int(hdr[1]) != n_chunks or (step is not None and int(hdr[2]) != step):
                raise RuntimeError(f"gather: rank {r} region not complete (magic {int(hdr[0]):#x}, chunks {int(hdr[1])}/{n_chunks}, "
                                   f"step {int(hdr[2])}, expected {step})")
            d = reg[8 * HDR_WORDS:8 * (HDR_WORDS + CH_WORDS * n_chunks)].view(np.int64).reshape(n_chunks, CH_WORDS)
            if not bool((d[:, CHUNK_FIELDS.index("done")] == 1).all()):
                raise RuntimeError(f"gather: rank {r} has unfinished chunks")
            self._dir.append(d)
            ii = np.asarray(idx, dtype=np.int64)
            pos = np.arange(len(ii))
            self.rank[ii] = r
            self.chunk[ii] = pos // lines_per_chunk
            self.k[ii] = pos % lines_per_chunk
        if (self.rank < 0).any():
            raise RuntimeError("gather: some lines are not covered by any shard")

    def __len__(self):
        return len(self.widths)

    def _f(self, r, c, name):
        return int(self._dir[r][c, CHUNK_FIELDS.index(name)])

    def _chunk_lines(self, r, c):
        key = (r, c)
        if key not in self._plan:
            idx = self.shards[r][c * self.lines_per_chunk:(c + 1) * self.lines_per_chunk]
            self._plan[key] = _lib.plan_lines([int(self.widths[i]) for i in idx])[0]
        return self._plan[key]

    def _loc(self, i):
        return int(self.rank[i]), int(self.chunk[i]), int(self.k[i])

    def mask(self, i: int) -> np.ndarray:
        """(128, W') u8 {0,255} glued + thresholded mask of input line i."""
        r, c, k = self._loc(i)
        ln = self._chunk_lines(r, c)[k]
        off = self._f(r, c, "off_planes") + int(ln["px_off"])
        pitch, w = int(ln["pitch"]), int(ln["width"])
        return self.arena.region(r)[off:off + 128 * pitch].reshape(128, pitch)[:, :w]

    def num(self, i: int) -> int:
        r, c, k = self._loc(i)
        off = self._f(r, c, "off_num")
        return int(self.arena.region(r)[off + 4 * k:off + 4 * k + 4].view(np.int32)[0])

    def _stat_off(self, r, c):
        n = self._f(r, c, "n_lines")
        off = self._f(r, c, "off_num")
        counts = self.arena.region(r)[off:off + 4 * n].view(np.int32).astype(np.int64) - 1
        return np.concatenate([[0], np.cumsum(counts)])

    def stats(self, i: int) -> np.ndarray:
        """cv2-layout stats rows (x, y, w, h, area) of the islands of line i, label order."""
        r, c, k = self._loc(i)
        so = self._stat_off(r, c)
        off = self._f(r, c, "off_stats")
        return self.arena.region(r)[off + 20 * int(so[k]):off + 20 * int(so[k + 1])].view(np.int32).reshape(-1, 5)

    def _group_range(self, r, c, k):
        n = self._f(r, c, "n_lines")
        off = self._f(r, c, "off_lgs")
        lgs = self.arena.region(r)[off:off + 8 * (n + 1)].view(np.int64)
        return int(lgs[k]), int(lgs[k + 1])

    def groups(self, i: int) -> np.ndarray:
        """(g, 6) int64 rows (line-in-chunk, left, top, right, bottom, canvas offset) in the reference's group order."""
        r, c, k = self._loc(i)
        a, b = self._group_range(r, c, k)
        off = self._f(r, c, "off_groups")
        return self.arena.region(r)[off + 48 * a:off + 48 * b].view(np.int64).reshape(-1, 6)

    def crops(self, i: int) -> np.ndarray:
        """(g, size, size) u8: the reference's partition `image` of every group of line i."""
        r, c, k = self._loc(i)
        a, b = self._group_range(r, c, k)
        size = self._f(r, c, "crop_size")
        off = self._f(r, c, "off_crops")
        return self.arena.region(r)[off + a * size * size:off + b * size * size].reshape(b - a, size, size)

    def canvases(self, i: int):
        """[(canvas u8 {0,1} (h, w), (top, left))] of line i (helper/partition.py:31-87)."""
        r, c, k = self._loc(i)
        off = self._f(r, c, "off_canvas")
        out = []
        for row in self.groups(i):
            _, left, top, right, bottom, o = (int(v) for v in row)
            h, w = bottom - top, right - left
            out.append((self.arena.region(r)[off + o:off + o + h * w].reshape(h, w), (np.int64(top), np.int64(left))))
        return out
