"""Device stages of the segmentation path and the batched line pipeline.

Everything here calls the C ABI (`libsd_b200.so`); torch only owns the device
buffers and streams.  Stage <-> reference map:
  tile_extract_*   helper/split.py:10-86 (+ evaluate_binarize.py:99)
  glue_*           helper/split.py:89-124 (+ evaluate_binarize.py:103, main.py:108)
  ccl_label        helper/partition.py:14   (cv2.connectedComponentsWithStats)
  island_stats     helper/partition.py:17-26
  group canvases   helper/partition.py:31-87 via the stats closed form
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib
from .engine import TILE_H, TILE_W, CIN_PAD, UNetEngine, stream_ptr

MARGIN = 2          # evaluate_strokes.py:26 / partition.py:9


@dataclass
class LineBatch:
    """A planned batch of text lines (all at height 128) resident on one GPU."""
    device: torch.device
    lines: np.ndarray                 # structured array, _lib.LINE_DTYPE
    plan: _lib.Plan
    d_lines: torch.Tensor             # uint8 view of the sd_line table on the device
    widths: list = field(default_factory=list)

    @property
    def n_lines(self): return int(self.plan.n_lines)
    @property
    def n_tiles(self): return int(self.plan.n_tiles)
    @property
    def px_total(self): return int(self.plan.px_total)
    @property
    def blk_total(self): return int(self.plan.blk_total)

    def plane(self, buf: torch.Tensor, i: int) -> torch.Tensor:
        """(128, W') view of line i inside a packed plane buffer."""
        ln = self.lines[i]
        off, pitch, w = int(ln["px_off"]), int(ln["pitch"]), int(ln["width"])
        return buf[off:off + TILE_H * pitch].view(TILE_H, pitch)[:, :w]

    def tile_dst(self, planes: torch.Tensor) -> torch.Tensor:
        """Device sd_tile_dst table (uint8 view, 16 B per tile) that pastes this batch's tiles into `planes`."""
        tab = _lib.tile_dst_table(self.lines, planes.data_ptr())
        return torch.from_numpy(tab.view(np.uint8).reshape(-1).copy()).to(self.device, non_blocking=True)

    # bookkeeping lists of cut_and_stack (helper/split.py:66-78)
    def stack_indices(self):
        return [list(range(int(l["first_tile"]), int(l["first_tile"] + l["n_tiles"]))) for l in self.lines]

    def stack_widths(self):
        out = []
        for l in self.lines:
            W, n, wu, ov = int(l["width"]), int(l["n_tiles"]), int(l["wu"]), int(l["overlap"])
            out.append([W] if n == 1 else [min((i + 1) * wu + ov, W) - i * wu for i in range(n)])
        return out


def plan_batch(widths, device, tile_w=TILE_W, overlap=64) -> LineBatch:
    lines, plan = _lib.plan_lines(widths, tile_w, overlap)
    raw = torch.from_numpy(lines.view(np.uint8).reshape(-1).copy())
    d_lines = raw.to(device, non_blocking=False)
    return LineBatch(torch.device(device), lines, plan, d_lines, [int(w) for w in widths])


def resized_width_hw(h: int, w: int, height: int = TILE_H) -> int:
    """Line width after resize_to_height (common.py:89-91): int(w * (height / h)); w itself at h == height,
    where cv2.resize is a copy."""
    return int(w) if h == height else int(w * (height / h))


def resized_width(img, height: int = TILE_H) -> int:
    return resized_width_hw(img.shape[0], img.shape[1], height)


def pack_lines_rgb(images, batch: LineBatch, pinned: bool = True, out: torch.Tensor | None = None) -> torch.Tensor:
    """Host-packs (128, W, 3) u8 images at sd_line.img_off (pinned staging; `out` = a staging tensor to fill).
    Lines of another height are left out: `ResizePlan` fills their slots on the device."""
    buf = out if out is not None else torch.empty(int(batch.plan.img_bytes), dtype=torch.uint8,
                                                  pin_memory=pinned and torch.cuda.is_available())
    assert buf.numel() >= int(batch.plan.img_bytes)
    nb = buf.numpy()

    def put(args):
        img, ln = args
        if img.shape[0] != TILE_H:
            return
        a = np.ascontiguousarray(img, dtype=np.uint8)
        assert a.shape == (TILE_H, int(ln["width"]), 3), (a.shape, int(ln["width"]))
        off = int(ln["img_off"])
        nb[off:off + a.size] = a.reshape(-1)
    work = list(zip(images, batch.lines))
    if int(batch.plan.img_bytes) >= (8 << 20):          # numpy copies release the GIL: a few threads beat one core's memcpy
        list(host_pool().map(put, work))
    else:
        for wk in work:
            put(wk)
    return buf


class ResizePlan:
    """The lines of a batch whose height is not 128 (resize_to_height, common.py:85-93): their pixels at the
    original size in one pinned buffer plus the sd_resize_job table; `run` resizes them on the device into
    their slots of the packed line buffer (sd_resize_lines, bit-exact with cv2.resize)."""

    def __init__(self, images, batch: LineBatch, pinned: bool = True):
        todo = [(i, im) for i, im in enumerate(images) if im.shape[0] != TILE_H]
        self.n = len(todo)
        self.device = batch.device
        if not self.n:
            return
        jobs = np.zeros(self.n, _lib.RESIZE_DTYPE)
        off = 0
        for j, (i, im) in enumerate(todo):
            if im.ndim != 3 or im.shape[2] != 3:
                raise ValueError(f"line {i}: expected an (h, w, 3) image, got {im.shape}")
            ln = batch.lines[i]
            jobs[j] = (off, int(ln["img_off"]), im.shape[0], im.shape[1], int(ln["width"]), 0)
            off += (im.size + 8 + 15) // 16 * 16        # the kernel's word loads may read 8 bytes past an image
        self.h_src = torch.empty(off, dtype=torch.uint8, pin_memory=pinned and torch.cuda.is_available())
        nb = self.h_src.numpy()
        for (i, im), jb in zip(todo, jobs):
            nb[int(jb["src_off"]):int(jb["src_off"]) + im.size] = np.ascontiguousarray(im, dtype=np.uint8).reshape(-1)
        self.max_dst_w = int(jobs["dst_w"].max())
        self.d_jobs = torch.from_numpy(jobs.view(np.uint8).reshape(-1).copy()).to(self.device)
        self.d_src = None

    def upload(self):
        if self.n:
            self.d_src = self.h_src.to(self.device, non_blocking=True) if self.d_src is None else self.d_src.copy_(
                self.h_src, non_blocking=True)

    def run(self, d_rgb: torch.Tensor):
        """Enqueues the resize on the current stream (after `upload` on the same stream or an event wait)."""
        if self.n:
            _lib.check(_lib.lib().sd_resize_lines(self.d_src.data_ptr(), self.d_jobs.data_ptr(), self.n, self.max_dst_w,
                                                  d_rgb.data_ptr(), stream_ptr(self.device)), "sd_resize_lines")


def upload_lines(images, device):
    """images of any height -> (batch, packed (128, W', 3) lines on the device): resize_to_height of every line
    (common.py:85-93; evaluate_binarize.py:76), with the lines that need it resized on the GPU."""
    batch = plan_batch([resized_width(im) for im in images], device)
    d_rgb = pack_lines_rgb(images, batch).to(device, non_blocking=True)
    rp = ResizePlan(images, batch)
    rp.upload()
    rp.run(d_rgb)
    return batch, d_rgb


def _s(batch): return stream_ptr(batch.device)


def tile_extract_f16(batch: LineBatch, d_rgb: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    if out is None:
        out = torch.empty((batch.n_tiles, TILE_H, TILE_W, CIN_PAD), dtype=_lib.torch_dtype(), device=batch.device)
    _lib.check(_lib.lib().sd_tile_extract_f16(d_rgb.data_ptr(), batch.d_lines.data_ptr(), batch.n_lines, batch.n_tiles,
                                              out.data_ptr(), _s(batch)), "sd_tile_extract_f16")
    return out


def tile_extract_u8(batch: LineBatch, d_rgb: torch.Tensor) -> torch.Tensor:
    out = torch.empty((batch.n_tiles, 3, TILE_H, TILE_W), dtype=torch.uint8, device=batch.device)
    _lib.check(_lib.lib().sd_tile_extract_u8(d_rgb.data_ptr(), batch.d_lines.data_ptr(), batch.n_lines, batch.n_tiles,
                                             out.data_ptr(), _s(batch)), "sd_tile_extract_u8")
    return out


def glue_u8(batch: LineBatch, tiles_u8: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    assert tiles_u8.dtype == torch.uint8 and tiles_u8.is_contiguous() and tiles_u8.numel() == batch.n_tiles * TILE_H * TILE_W
    if out is None:
        out = torch.empty(batch.px_total, dtype=torch.uint8, device=batch.device)
    _lib.check(_lib.lib().sd_glue_u8(tiles_u8.data_ptr(), batch.n_tiles, batch.d_lines.data_ptr(), batch.n_lines,
                                     batch.px_total, out.data_ptr(), _s(batch)), "sd_glue_u8")
    return out


def glue_threshold_f16(batch: LineBatch, prob16: torch.Tensor, bin_thr=0.5, on_value=255) -> torch.Tensor:
    assert prob16.dtype == torch.float16 and prob16.is_contiguous() and prob16.numel() == batch.n_tiles * TILE_H * TILE_W
    out = torch.empty(batch.px_total, dtype=torch.uint8, device=batch.device)
    _lib.check(_lib.lib().sd_glue_threshold_f16(prob16.data_ptr(), batch.n_tiles, batch.d_lines.data_ptr(), batch.n_lines,
                                                batch.px_total, float(bin_thr), int(on_value), out.data_ptr(), _s(batch)),
               "sd_glue_threshold_f16")
    return out


def ccl_label(batch: LineBatch, planes: torch.Tensor, work: torch.Tensor | None = None):
    """-> (labels int32 packed planes, num int32[n_lines] incl. background)."""
    L = _lib.lib()
    need = L.sd_ccl_workspace_bytes(batch.blk_total, batch.n_lines)
    if work is None or work.numel() < need:
        work = torch.empty(need, dtype=torch.uint8, device=batch.device)
    labels = torch.empty(batch.px_total, dtype=torch.int32, device=batch.device)
    num = torch.empty(batch.n_lines, dtype=torch.int32, device=batch.device)
    _lib.check(L.sd_ccl_label(planes.data_ptr(), batch.d_lines.data_ptr(), batch.n_lines, batch.px_total, batch.blk_total,
                              labels.data_ptr(), num.data_ptr(), work.data_ptr(), _s(batch)), "sd_ccl_label")
    return labels, num


def stats_capacity(batch: LineBatch) -> int:
    """Rows reserved for the fused island statistics: ~6x the island density of handwriting-like lines
    (~20 islands per 1000 px of line width); a batch that needs more is re-run with the exact count."""
    return int(batch.px_total // (TILE_H * 8)) + 1024


def ccl_label_stats(batch: LineBatch, planes: torch.Tensor, cap_rows: int, work: torch.Tensor | None = None):
    """Labels + cv2-layout island stats in one pass (sd_ccl_label_stats).
    -> (labels int32 packed planes, meta uint8 = [stat_off int64 (n_lines+1) | num int32 (n_lines)], stats int32 (cap_rows, 5))."""
    L = _lib.lib()
    need = L.sd_ccl_workspace_bytes(batch.blk_total, batch.n_lines)
    if work is None or work.numel() < need:
        work = torch.empty(need, dtype=torch.uint8, device=batch.device)
    n = batch.n_lines
    labels = torch.empty(batch.px_total, dtype=torch.int32, device=batch.device)
    meta = torch.empty(8 * (n + 1) + 4 * n, dtype=torch.uint8, device=batch.device)
    stats = torch.empty((cap_rows, 5), dtype=torch.int32, device=batch.device)
    _lib.check(L.sd_ccl_label_stats(planes.data_ptr(), batch.d_lines.data_ptr(), n, batch.px_total, batch.blk_total,
                                    labels.data_ptr(), meta.data_ptr() + 8 * (n + 1), meta.data_ptr(), stats.data_ptr(), cap_rows,
                                    work.data_ptr(), _s(batch)), "sd_ccl_label_stats")
    return labels, meta, stats


def island_stats(batch: LineBatch, labels: torch.Tensor, num_host: np.ndarray):
    """-> (stats int32 (rows,5) = x,y,w,h,area per label, stat_off int64[n_lines+1] host, d_stat_off)."""
    counts = np.asarray(num_host, dtype=np.int64) - 1
    stat_off = np.zeros(batch.n_lines + 1, np.int64)
    np.cumsum(counts, out=stat_off[1:])
    rows = int(stat_off[-1])
    d_off = torch.from_numpy(stat_off[:-1].copy()).to(batch.device)
    stats = torch.empty((max(rows, 1), 5), dtype=torch.int32, device=batch.device)
    _lib.check(_lib.lib().sd_island_stats(labels.data_ptr(), batch.d_lines.data_ptr(), batch.n_lines, batch.px_total,
                                          d_off.data_ptr(), rows, stats.data_ptr(), _s(batch)), "sd_island_stats")
    return stats[:rows], stat_off, d_off


def island_boxes(stats_line: np.ndarray, W: int, H: int = TILE_H, margin: int = MARGIN):
    """helper/partition.py:19-24: margin-expanded boxes (xs, ys, xf, yf)."""
    x, y, w, h = (stats_line[:, k].astype(np.int64) for k in range(4))
    xs = np.maximum(x - margin, 0); ys = np.maximum(y - margin, 0)
    xf = np.minimum(x + w + margin + 1, W); yf = np.minimum(y + h + margin + 1, H)
    return xs, ys, xf, yf


def group_line(stats_line: np.ndarray, W: int, target_w: int = TILE_H, margin: int = MARGIN):
    """Grouping half of group_islands (helper/partition.py:38-69) from stats.
    Returns (groups: list of member label arrays (1-based), boxes int64 (g,4) = left, top, right, bottom)."""
    n = len(stats_line)
    if n == 0:
        return [], np.zeros((0, 4), np.int64)
    xs, ys, xf, yf = island_boxes(stats_line, W, margin=margin)
    # sort_islands (:90-98): the SAME np.argsort call on a Python list of ints (tie order)
    order = np.argsort([int(v) for v in xs])
    iv = np.stack([xs[order], xf[order]], axis=1)
    groups = _lib.group_intervals(iv, target_w)
    out_groups, boxes = [], np.zeros((len(groups), 4), np.int64)
    for g, members in enumerate(groups):
        idx = order[np.asarray(members, dtype=np.int64)]
        boxes[g] = (xs[idx].min(), ys[idx].min(), xf[idx].max(), yf[idx].max())
        out_groups.append(idx + 1)
    return out_groups, boxes


def group_canvases(batch: LineBatch, labels: torch.Tensor, stat_off: np.ndarray, d_stat_off: torch.Tensor,
                   line_groups, line_boxes):
    """Device canvases for every group of every line.
    line_groups[l] = list of member-label arrays, line_boxes[l] = (g,4) boxes.
    Returns list per line of [(canvas u8 (h,w), (top, left))]."""
    rows = int(stat_off[-1])
    n_groups = sum(len(g) for g in line_groups)
    if n_groups == 0:
        return [[] for _ in line_groups]
    table = np.zeros((n_groups, 6), np.int64)
    group_of = np.full(max(rows, 1), -1, np.int32)
    g = 0
    off = 0
    for l, (groups, boxes) in enumerate(zip(line_groups, line_boxes)):
        for members, (left, top, right, bottom) in zip(groups, boxes):
            table[g] = (l, left, top, right, bottom, off)
            group_of[stat_off[l] + members - 1] = g
            off += int((right - left) * (bottom - top))
            g += 1
    d_table = torch.from_numpy(table).to(batch.device)
    d_gof = torch.from_numpy(group_of).to(batch.device)
    canvas = torch.empty(max(off, 1), dtype=torch.uint8, device=batch.device)
    _lib.check(_lib.lib().sd_group_canvas(labels.data_ptr(), batch.d_lines.data_ptr(), d_table.data_ptr(), n_groups,
                                          d_gof.data_ptr(), d_stat_off.data_ptr(), canvas.data_ptr(), _s(batch)),
               "sd_group_canvas")
    host = canvas.cpu().numpy()
    out = [[] for _ in line_groups]
    for row in table:
        l, left, top, right, bottom, o = (int(v) for v in row)
        h, w = bottom - top, right - left
        out[l].append((host[o:o + h * w].reshape(h, w), (np.int64(top), np.int64(left))))
    return out


IMG_SIZE = 224      # evaluate_strokes.py:25
IMAGENET_MEAN = [0.485, 0.456, 0.406]   # evaluate_strokes.py:27-28
IMAGENET_STD = [0.229, 0.224, 0.225]
CROP_MARGIN = 1     # evaluate_strokes.py:207 (resize_and_pad_image(..., margin=1))


def crop_geometry(groups: np.ndarray, size: int = IMG_SIZE, margin: int = CROP_MARGIN):
    """helper/partition.py:112-139 for every row of the group table, vectorised in float64 (the same IEEE
    operations as the reference's Python floats).  -> (rs_dims int32 (g,2) = rs_w, rs_h; ratio f64 (g,);
    translate2 f64 (g,2) = x, y)."""
    w = (groups[:, 3] - groups[:, 1]).astype(np.float64)
    h = (groups[:, 4] - groups[:, 2]).astype(np.float64)
    new = float(size - 2 * margin)
    scale = np.minimum(new / h, new / w)
    rs_w = np.minimum(np.rint(scale * w), new).astype(np.int64)
    rs_h = np.minimum(np.rint(scale * h), new).astype(np.int64)
    ratio = (rs_w / w + rs_h / h) / 2
    t2 = np.stack([(size - rs_w) / 2, (size - rs_h) / 2], axis=1)
    return np.stack([rs_w, rs_h], axis=1).astype(np.int32), ratio, t2


def input_lut(mean, std) -> np.ndarray:
    """(3, 256) f32 table of evaluate_strokes.py:66-68: ((v / 255. - mean[c]) / std[c]).astype(float32)."""
    v = np.arange(256, dtype=np.uint8)
    return np.stack([(v / 255. - mean[c]) / std[c] for c in range(3)], axis=0).astype(np.float32)


def group_crops(device, canvas: torch.Tensor, d_groups: torch.Tensor, groups: np.ndarray, size: int = IMG_SIZE,
                margin: int = CROP_MARGIN, lut: np.ndarray | None = None):
    """Device crops of every group: -> dict(image u8 (g,size,size), image_input f32 (g,3,size,size) | None,
    ratio, translate2, rs_dims)."""
    n = len(groups)
    rs, ratio, t2 = crop_geometry(groups, size, margin)
    image = torch.empty((n, size, size), dtype=torch.uint8, device=device)
    inp = d_lut = None
    if lut is not None:
        inp = torch.empty((n, 3, size, size), dtype=torch.float32, device=device)
        d_lut = torch.from_numpy(np.ascontiguousarray(lut, np.float32)).to(device)
    if n:
        d_rs = torch.from_numpy(rs).to(device)
        _lib.check(_lib.lib().sd_group_crops(canvas.data_ptr(), d_groups.data_ptr(), d_rs.data_ptr(), n, size, image.data_ptr(),
                                             inp.data_ptr() if inp is not None else None,
                                             d_lut.data_ptr() if d_lut is not None else None,
                                             torch.cuda.current_stream(device).cuda_stream), "sd_group_crops")
    return {"image": image, "image_input": inp, "ratio": ratio, "translate2": t2, "rs_dims": rs}


_POOL = None


def host_pool():
    """A few host threads for the big numpy copies of the batched calls (packing lines, fresh copies of results)."""
    global _POOL
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(max_workers=4, thread_name_prefix="sd-host")
    return _POOL


_LANES = None


def lane_pool():
    """Threads of the host lanes of the batched calls (a lane = a thread + a CUDA stream that takes every other chunk)."""
    global _LANES
    if _LANES is None:
        from concurrent.futures import ThreadPoolExecutor
        _LANES = ThreadPoolExecutor(max_workers=4, thread_name_prefix="sd-lane")
    return _LANES


class PinnedStaging:
    """Grow-only pinned host staging buffers, one per purpose (`key`), owned by ONE Segmenter / job (two
    Segmenters driven from different threads, one per GPU, never share a buffer).  `get` returns a uint8 numpy
    view; device-to-host copies into it go through sd_copy_d2h_async."""

    def __init__(self):
        self._bufs = {}

    def get_tensor(self, key, nbytes: int) -> torch.Tensor:
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(int(nbytes * 1.25), 1 << 16), dtype=torch.uint8, pin_memory=torch.cuda.is_available())
            self._bufs[key] = buf
        return buf[:nbytes]

    def get(self, key, nbytes: int) -> np.ndarray:
        return self.get_tensor(key, nbytes).numpy()


class FreshPinned:
    """Staging for RESULTS the caller keeps: every `get` hands out a new page-locked buffer, so device-to-host copies
    land directly in the arrays that are returned (no second host copy).  The numpy view keeps its tensor alive; when
    the last view of a result dies the block goes back to torch's caching host allocator, and the next call gets it
    without a cudaHostAlloc (first calls pay the page-locking once)."""

    def get_tensor(self, key, nbytes: int) -> torch.Tensor:
        return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, pin_memory=torch.cuda.is_available())[:nbytes]

    def get(self, key, nbytes: int) -> np.ndarray:
        return self.get_tensor(key, nbytes).numpy()


def copy_d2h(dst: np.ndarray, src: torch.Tensor, device) -> np.ndarray:
    """Asynchronous device -> (page-locked) host copy on the current stream of `device`; returns `dst`."""
    assert dst.flags.c_contiguous and src.is_contiguous() and dst.nbytes == src.numel() * src.element_size(), (dst.nbytes, src.shape)
    _lib.check(_lib.lib().sd_copy_d2h_async(dst.ctypes.data, src.data_ptr(), dst.nbytes, stream_ptr(device)), "sd_copy_d2h_async")
    return dst


class LazyPartition(dict):
    """One partition dict of get_partitions (evaluate_strokes.py:213-219): `image`, `translate1`, `ratio`,
    `translate2` are stored; `image_input` ((3, size, size) f32, 12x the bytes of `image`) is built on first
    `part["image_input"]` from `image` exactly like `_normalize_image` (:58-69: MINMAX to 0..255, then per channel
    (v / 255. - mean) / std through the same float64 table).  On the device path the f32 tensor of ALL crops is
    available without crossing PCIe (`PartitionResult["crops"]["image_input"]`)."""

    def __init__(self, lut, **kw):
        super().__init__(**kw)
        self._lut = lut

    def __contains__(self, key):
        return key == "image_input" or dict.__contains__(self, key)

    def get(self, key, default=None):
        return self[key] if key in self else default

    def __missing__(self, key):
        if key != "image_input":
            raise KeyError(key)
        import cv2
        norm = cv2.normalize(np.ascontiguousarray(self["image"]), None, 0, 255, norm_type=cv2.NORM_MINMAX)
        v = self._lut[:, norm]
        self[key] = v
        return v


_LP_CLASSES = {}


def lazy_partition_class(lut: np.ndarray):
    """A LazyPartition subclass with the table bound at class level: instances are then built by dict's own C
    constructor (`cls(image=..., translate1=..., ...)`), ~3x cheaper than a Python __init__ per partition."""
    key = lut.tobytes()
    cls = _LP_CLASSES.get(key)
    if cls is None:
        cls = type("LazyPartition", (LazyPartition,), {"_lut": lut, "__init__": dict.__init__})
        if len(_LP_CLASSES) > 16:
            _LP_CLASSES.clear()
        _LP_CLASSES[key] = cls
    return cls


def build_partitions(lut, image_host, groups, crops, lgs, n_lines):
    """Per line the list of partition dicts of evaluate_strokes.py:213-219 for one chunk: `image` is a (size, size) view
    into `image_host`, translate1 = (left, top) as numpy int64 like the reference's, ratio / translate2 Python floats."""
    if image_host is None or not len(groups):
        return [[] for _ in range(n_lines)]
    LP = lazy_partition_class(lut)
    t1 = zip(groups[:, 1], groups[:, 2])
    t2 = zip(crops["translate2"][:, 0].tolist(), crops["translate2"][:, 1].tolist())
    flat = [LP(image=im, translate1=a, ratio=r, translate2=b) for im, a, r, b in zip(image_host, t1, crops["ratio"].tolist(), t2)]
    b = lgs.tolist()
    return [flat[b[k]:b[k + 1]] for k in range(n_lines)]


class PartitionResult(dict):
    """Result of `Segmenter.partition`: labels / canvas stay on the device, tables on the host."""

    def line_canvases(self, l: int):
        """[(canvas u8 {0,1} (h,w), (top, left))] of line l, in the reference's group order."""
        host = self["canvas_host"]
        if host is None:
            host = self["canvas"].cpu().numpy()
            self["canvas_host"] = host
        a, b = int(self["line_group_start"][l]), int(self["line_group_start"][l + 1])
        out = []
        for row in self["groups"][a:b]:
            _, left, top, right, bottom, o = (int(v) for v in row)
            h, w = bottom - top, right - left
            out.append((host[o:o + h * w].reshape(h, w), (np.int64(top), np.int64(left))))
        return out

    @property
    def canvases(self):
        return [self.line_canvases(l) for l in range(len(self["line_group_start"]) - 1)]

    def line_partitions(self, l: int):
        """evaluate_strokes.py:202-222 for line l from the device crops (needs partition(..., crops=...)):
        [dict(image, image_input, translate1=(left, top), ratio, translate2)]."""
        cr = self["crops"]
        if cr is None:
            raise RuntimeError("partition() was called without crops")
        if "image_host" not in cr:
            cr["image_host"] = cr["image"].cpu().numpy()
            cr["input_host"] = cr["image_input"].cpu().numpy() if cr["image_input"] is not None else None
        a, b = int(self["line_group_start"][l]), int(self["line_group_start"][l + 1])
        out = []
        for g in range(a, b):
            _, left, top, _, _, _ = (int(v) for v in self["groups"][g])
            out.append({"image": cr["image_host"][g],
                        "image_input": cr["input_host"][g] if cr["input_host"] is not None else None,
                        "translate1": (np.int64(left), np.int64(top)), "ratio": float(cr["ratio"][g]),
                        "translate2": (float(cr["translate2"][g, 0]), float(cr["translate2"][g, 1]))})
        return out


class Segmenter:
    """Batched text segmentation of many line images on ONE GPU:
    tile -> Attention-UNet -> glue/threshold -> CCL -> island boxes -> group canvases.

    Equivalent to, per line, `BinarizationSession.binarize_image` + `main.py:108`
    + `get_binarized_islands` + `group_islands` of the reference, but every stage
    runs once over the whole batch."""

    SPEC_ROWS = 16384       # stats rows copied to the host together with the counts (one sync for typical chunks)

    def __init__(self, engine: UNetEngine | None, bin_thr: float = 0.5, margin: int = MARGIN, device=None):
        self.engine = engine
        self.device = engine.device if engine is not None else torch.device(device)
        self.bin_thr = bin_thr
        self.margin = margin
        self.staging = PinnedStaging()
        self._streams = None
        self._rows_per_px = 0.0          # densest batch seen so far (islands per pixel): sizes the stats reserve of the next one

    @classmethod
    def for_engine(cls, engine: UNetEngine, bin_thr: float = 0.5) -> "Segmenter":
        """The Segmenter kept with an engine across calls: its pinned staging buffers and its CUDA streams are
        reused (the caching allocator pools device memory per stream: fresh streams per call would turn every
        allocation of a call into a cudaMalloc)."""
        seg = getattr(engine, "_sd_segmenter", None)
        if seg is None or seg.bin_thr != bin_thr:
            seg = cls(engine, bin_thr=bin_thr)
            engine._sd_segmenter = seg
        return seg

    def streams(self):
        """(copy, unet, part) streams of the chunk pipeline, created once per Segmenter."""
        if self._streams is None:
            with torch.cuda.device(self.device):
                self._streams = tuple(torch.cuda.Stream(self.device) for _ in range(3))
        return self._streams

    def lane_streams(self, n: int):
        """CUDA streams of the host lanes (`lane_pool`), created once per Segmenter."""
        if getattr(self, "_lane_streams", None) is None or len(self._lane_streams) < n:
            with torch.cuda.device(self.device):
                self._lane_streams = tuple(torch.cuda.Stream(self.device) for _ in range(n))
        return self._lane_streams[:n]

    def binarize(self, images, d_rgb: torch.Tensor | None = None, batch: LineBatch | None = None):
        """-> (batch, mask planes u8 {0,255} packed on device)."""
        with torch.cuda.device(self.device):
            if batch is None and d_rgb is None:
                batch, d_rgb = upload_lines(images, self.device)       # any height: resize_to_height on the GPU
            elif batch is None:
                batch = plan_batch([im.shape[1] for im in images], self.device)
            elif d_rgb is None:
                d_rgb = pack_lines_rgb(images, batch).to(self.device, non_blocking=True)
            tiles = tile_extract_f16(batch, d_rgb)
            # reconstruct_images (helper/split.py:89-124) is fused into the UNet head: tiles OR their thresholded
            # columns straight into the zeroed line planes, no tile-shaped mask and no glue pass exist
            planes = torch.zeros(batch.px_total, dtype=torch.uint8, device=self.device)
            dst = batch.tile_dst(planes)
            mt = self.engine.max_tiles
            for s in range(0, batch.n_tiles, mt):
                self.engine.forward_lines(tiles[s:s + mt], dst[16 * s:16 * (s + mt)], self.bin_thr)
        return batch, planes

    def partition(self, batch: LineBatch, planes: torch.Tensor, canvases: str = "host", key="part",
                  zero_copy: bool = False, crops: bool = False, crop_lut: np.ndarray | None = None,
                  staging=None, crops_to_host: bool = False) -> PartitionResult:
        """mask planes -> labels, island stats, groups and group canvases for every line.
        canvases: "host" (copied to page-locked host memory), "device" (left in HBM) or "none".
        staging: where host copies land (`get(key, nbytes) -> uint8 numpy view`): this Segmenter's pinned buffers
        by default, a gather arena region for multi-GPU jobs.  zero_copy: host arrays alias the staging memory
        (valid until the next call with the same key) instead of being copied out.
        crops_to_host: also copy the u8 crops (`image`) to the staging memory."""
        st = staging if staging is not None else self.staging
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device)
            n = batch.n_lines
            # labels and the cv2 stats rows in ONE pass (no second read of the labels); counts, row offsets and a
            # first slice of the rows come back in one D2H + one sync
            # the reserve follows the densest batch seen so far (+50 %): the masks of one job look alike, so only the
            # first dense batch pays the second run
            want = int(1.5 * self._rows_per_px * batch.px_total) + 1024
            cap = max(stats_capacity(batch), want)
            labels, meta, stats = ccl_label_stats(batch, planes, cap)
            spec = min(cap, max(self.SPEC_ROWS, want))
            h_meta = copy_d2h(self.staging.get((key, "meta"), meta.numel()), meta, self.device)
            h_spec = copy_d2h(self.staging.get((key, "spec"), spec * 20), stats[:spec], self.device)
            stream.synchronize()
            stat_off = h_meta[:8 * (n + 1)].view(np.int64).copy()
            rows = int(stat_off[-1])
            self._rows_per_px = max(self._rows_per_px, rows / max(batch.px_total, 1))
            if rows > cap:                            # denser than the reserve: once more with the exact room
                labels, meta, stats = ccl_label_stats(batch, planes, rows)
                cap, spec = rows, 0
            h_num = st.get((key, "num"), n * 4).view(np.int32)
            h_num[:] = h_meta[8 * (n + 1):].view(np.int32)
            h_stats = st.get((key, "stats"), rows * 20).view(np.int32).reshape(rows, 5)
            if rows <= spec:
                h_stats[:] = h_spec.view(np.int32).reshape(-1, 5)[:rows]
            elif rows:
                copy_d2h(h_stats, stats[:rows], self.device)
                stream.synchronize()
            d_off = meta[:8 * n].view(torch.int64)
            num_h = h_num if zero_copy else h_num.copy()
            stats_h = h_stats if zero_copy else h_stats.copy()
            groups, group_of, lgs, cbytes = _lib.group_lines(stats_h, stat_off, batch.widths, self.margin, TILE_H, TILE_H)
            res = PartitionResult(labels=labels, num=num_h, stats=stats_h, stat_off=stat_off, groups=groups,
                                  group_of=group_of, line_group_start=lgs, canvas=None, canvas_host=None,
                                  canvas_bytes=cbytes, crops=None)
            if canvases != "none" and len(groups):
                d_table = torch.from_numpy(groups).to(self.device, non_blocking=True)
                d_gof = torch.from_numpy(group_of).to(self.device, non_blocking=True)
                canvas = torch.empty(max(cbytes, 1), dtype=torch.uint8, device=self.device)
                _lib.check(_lib.lib().sd_group_canvas(labels.data_ptr(), batch.d_lines.data_ptr(), d_table.data_ptr(),
                                                      len(groups), d_gof.data_ptr(), d_off.data_ptr(), canvas.data_ptr(),
                                                      _s(batch)), "sd_group_canvas")
                res["canvas"] = canvas[:cbytes]
                res["_keep"] = (d_table, d_gof, d_off)
                if crops:     # 224x224 stroke-estimator crops straight from the device canvases
                    res["crops"] = group_crops(self.device, canvas, d_table, groups, lut=crop_lut)
                    if crops_to_host:
                        img = res["crops"]["image"]
                        res["crops"]["image_host"] = copy_d2h(st.get((key, "crops"), img.numel()), img, self.device).reshape(tuple(img.shape))
                        res["crops"]["input_host"] = None
                if canvases == "host":
                    hb = copy_d2h(st.get((key, "canvas"), cbytes), res["canvas"], self.device)
                    stream.synchronize()
                    res["canvas_host"] = hb if zero_copy else hb.copy()
            elif canvases != "none":
                res["canvas"] = torch.empty(0, dtype=torch.uint8, device=self.device)
                res["canvas_host"] = np.zeros(0, np.uint8)
                if crops:
                    res["crops"] = group_crops(self.device, res["canvas"], None, groups, lut=crop_lut)
        return res

    def segment(self, images):
        batch, planes = self.binarize(images)
        res = self.partition(batch, planes)
        res["batch"] = batch
        res["planes"] = planes
        return res
