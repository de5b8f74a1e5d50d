"""ctypes binding of `libsd_b200.so` (the C ABI declared in include/sd_b200.h).

There is no CPU fallback: if the library is missing it is built (nvcc), and if
that fails, or a compute entry point is called without a CUDA device, the call
raises.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import build as _build

HERE = Path(__file__).resolve().parent

# numpy mirror of `struct sd_line` (include/sd_b200.h)
LINE_DTYPE = np.dtype([
    ("img_off", "<i8"), ("px_off", "<i8"), ("blk_off", "<i8"),
    ("width", "<i4"), ("n_tiles", "<i4"), ("wu", "<i4"), ("first_tile", "<i4"),
    ("tile_w", "<i4"), ("overlap", "<i4"), ("pitch", "<i4"), ("bw", "<i4"),
], align=True)
assert LINE_DTYPE.itemsize == 56

# numpy mirror of `struct sd_resize_job`
RESIZE_DTYPE = np.dtype([("src_off", "<i8"), ("dst_off", "<i8"), ("src_h", "<i4"), ("src_w", "<i4"), ("dst_w", "<i4"),
                         ("reserved", "<i4")], align=True)
assert RESIZE_DTYPE.itemsize == 32


class Plan(C.Structure):
    _fields_ = [("img_bytes", C.c_int64), ("px_total", C.c_int64), ("blk_total", C.c_int64),
                ("n_tiles", C.c_int32), ("n_lines", C.c_int32)]


SLOTS = [
    "CONV1_0", "CONV1_1", "CONV2_0", "CONV2_1", "CONV3_0", "CONV3_1", "CONV4_0", "CONV4_1", "CONV5_0", "CONV5_1",
    "UP5", "ATT5_G", "ATT5_X", "ATT5_PSI", "UPCONV5_0", "UPCONV5_1",
    "UP4", "ATT4_G", "ATT4_X", "ATT4_PSI", "UPCONV4_0", "UPCONV4_1",
    "UP3", "ATT3_G", "ATT3_X", "ATT3_PSI", "UPCONV3_0", "UPCONV3_1",
    "UP2", "ATT2_G", "ATT2_X", "ATT2_PSI", "UPCONV2_0", "UPCONV2_1",
    "HEAD",
]
TAPS = ["x1", "x2", "x3", "x4", "x5", "d5u", "a4", "d5", "d4u", "a3", "d4", "d3u", "a2", "d3", "d2u", "a1", "d2"]

EXPORTS = {
    # name: (restype, argtypes)
    "sd_last_error": (C.c_char_p, []),
    "sd_version": (C.c_int, []),
    "sd_operand_dtype": (C.c_char_p, []),
    "sd_cuda_available": (C.c_int, []),
    "sd_launch_count": (C.c_int64, []),
    "sd_plan_lines": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(Plan)]),
    "sd_tile_dst_table": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "sd_group_intervals": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    "sd_group_lines": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int64,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sd_tile_extract_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "sd_tile_extract_f16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "sd_glue_u8": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    "sd_glue_threshold_f16": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_int,
                                        C.c_void_p, C.c_void_p]),
    "sd_ccl_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int]),
    "sd_ccl_label": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p]),
    "sd_ccl_label_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "sd_island_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                  C.c_void_p]),
    "sd_group_canvas": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p]),
    "sd_group_crops": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p]),
    "sd_encode_postprocess": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "sd_resize_lines": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "sd_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "sd_host_unregister": (C.c_int, [C.c_void_p]),
    "sd_copy_d2h_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "sd_engine_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "sd_engine_destroy": (None, [C.c_void_p]),
    "sd_engine_set_conv": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "sd_engine_finalize": (C.c_int, [C.c_void_p, C.c_int]),
    "sd_engine_set_head_bias": (C.c_int, [C.c_void_p, C.c_float]),
    "sd_unet_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p]),
    "sd_unet_forward_lines": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "sd_unet_read_tap": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_int),
                                   C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p]),
    "sd_engine_enable_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "sd_engine_layer_times": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "sd_engine_layer_name": (C.c_char_p, [C.c_void_p, C.c_int]),
    "sd_debug_wait_cycles": (C.c_int, [C.c_void_p, C.c_int]),
    "sd_engine_wait_error": (C.c_int, [C.c_void_p]),
}


class SdError(RuntimeError):
    pass


_lib = None


def operand_dtype() -> str:
    """Operand type of the UNet in this process: "f16" (default, the parity build) or "bf16" (SD_DTYPE=bf16 selects
    libsd_b200_bf16.so, the same sources built with -DSD_BF16)."""
    import os
    dt = os.environ.get("SD_DTYPE", "f16").lower()
    return {"fp16": "f16", "half": "f16", "bfloat16": "bf16"}.get(dt, dt)


def torch_dtype():
    import torch
    return torch.bfloat16 if operand_dtype() == "bf16" else torch.float16


def lib() -> C.CDLL:
    """Loads (building first if needed) the native library; raises if impossible."""
    global _lib
    if _lib is None:
        path = _build.build(dtype=operand_dtype())   # no-op when the in-tree library matches the sources (digest stamp)
        handle = C.CDLL(str(path))
        for name, (res, args) in EXPORTS.items():
            fn = getattr(handle, name)      # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().sd_last_error().decode(errors="replace")
        raise SdError(f"{what or 'sd_b200'} failed ({rc}): {msg}")


def require_cuda() -> None:
    if not lib().sd_cuda_available():
        raise SdError("libsd_b200: no CUDA device visible; this path has no CPU fallback")


def plan_lines(widths, tile_w: int = 384, overlap: int = 64):
    """-> (lines structured array, Plan). Host-only (sd_plan_lines)."""
    w = np.ascontiguousarray(widths, dtype=np.int32)
    lines = np.zeros(len(w), dtype=LINE_DTYPE)
    plan = Plan()
    check(lib().sd_plan_lines(w.ctypes.data, len(w), tile_w, overlap, lines.ctypes.data, C.byref(plan)), "sd_plan_lines")
    return lines, plan


# numpy mirror of `struct sd_tile_dst`
TILE_DST_DTYPE = np.dtype([("d_dst", "<u8"), ("pitch", "<i4"), ("width", "<i4")], align=True)
assert TILE_DST_DTYPE.itemsize == 16


def tile_dst_table(lines, planes_ptr: int):
    """-> structured array (n_tiles,) of sd_tile_dst for planes starting at device address `planes_ptr`."""
    n_tiles = int(lines["n_tiles"].sum()) if len(lines) else 0
    out = np.zeros(max(n_tiles, 1), dtype=TILE_DST_DTYPE)
    lines = np.ascontiguousarray(lines)
    check(lib().sd_tile_dst_table(lines.ctypes.data, len(lines), planes_ptr, out.ctypes.data), "sd_tile_dst_table")
    return out[:n_tiles]


def group_intervals(intervals, width: int):
    """Native restatement of helper/partition.py:248-358 -> list[list[int]]."""
    iv = np.ascontiguousarray(np.asarray(intervals, dtype=np.int64).reshape(-1, 2))
    n = len(iv)
    members = np.zeros(max(n, 1), np.int32)
    starts = np.zeros(n + 1, np.int32)
    ng = lib().sd_group_intervals(iv.ctypes.data, n, int(width), members.ctypes.data, starts.ctypes.data)
    if ng < 0:
        check(ng, "sd_group_intervals")
    return [members[starts[g]:starts[g + 1]].tolist() for g in range(ng)]


def group_lines(stats, stat_off, widths, margin: int = 2, img_h: int = 128, target_w: int = 128):
    """Batched clustering of every line's islands (sd_group_lines).
    stats: (rows,5) int32 cv2 layout; stat_off: int64[n_lines+1]; widths: per line.
    -> (groups int64 (n_groups,6), group_of int32 (rows,), line_group_start int64[n_lines+1], canvas_bytes)."""
    stats = np.ascontiguousarray(stats, dtype=np.int32).reshape(-1, 5)
    stat_off = np.ascontiguousarray(stat_off, dtype=np.int64)
    widths = np.ascontiguousarray(widths, dtype=np.int32)
    rows, n_lines = len(stats), len(widths)
    # sort_islands (helper/partition.py:90-98): np.argsort of the left edges, per line, default kind
    xs = np.maximum(stats[:, 0].astype(np.int64) - margin, 0)
    order = np.empty(rows, np.int64)
    for l in range(n_lines):
        a, b = stat_off[l], stat_off[l + 1]
        if b > a:
            order[a:b] = np.argsort(xs[a:b])
    groups = np.zeros((max(rows, 1), 6), np.int64)
    group_of = np.full(max(rows, 1), -1, np.int32)
    lgs = np.zeros(n_lines + 1, np.int64)
    cbytes = C.c_int64()
    ng = lib().sd_group_lines(stats.ctypes.data, stat_off.ctypes.data, widths.ctypes.data, n_lines, order.ctypes.data,
                              margin, img_h, target_w, groups.ctypes.data, group_of.ctypes.data, lgs.ctypes.data,
                              C.byref(cbytes))
    if ng < 0:
        check(int(ng), "sd_group_lines")
    return groups[:ng], group_of[:rows], lgs, int(cbytes.value)
