"""GPU-backed mirror of the island half of
/root/reference/derenderer/helper/partition.py (get_binarized_islands,
group_islands, sort_islands, group_intervals, resize_and_pad_image, get_pad_edges).
Labelling, island boxes and crops run on the GPU (sd_ccl_label, sd_island_stats,
sd_group_canvas); interval grouping is the native host routine sd_group_intervals.
The stroke post-processing helpers of that file are out of scope.
"""

import cv2
import numpy as np
import torch

from .. import _lib
from .. import segment as _seg

_DEVICE = 0


def set_device(device: int):
    global _DEVICE
    _DEVICE = int(device)


def _label_image(img_bin):
    """-> (batch, labels device tensor, num, stats (N,5) host, stat_off, d_off)."""
    img = np.ascontiguousarray(img_bin)
    if img.ndim != 2 or img.shape[0] != _seg.TILE_H:
        raise ValueError(f"B200 path labels (128, W) masks; got {img.shape}")
    dev = torch.device("cuda", _DEVICE)
    batch = _seg.plan_batch([img.shape[1]], dev)
    pitch = int(batch.lines[0]["pitch"])
    host = np.zeros((_seg.TILE_H, pitch), np.uint8)
    host[:, :img.shape[1]] = img != 0
    planes = torch.from_numpy(host.reshape(-1)).to(dev)
    labels, num = _seg.ccl_label(batch, planes)
    num_h = num.cpu().numpy()
    stats, stat_off, d_off = _seg.island_stats(batch, labels, num_h)
    return batch, labels, int(num_h[0]), stats.cpu().numpy(), stat_off, d_off


def get_binarized_islands(img_bin, margin=2):
    """partition.py:9-28 -> (islands [(crop u8 {0,1}, (ys, xs))], img_islands int32, num_islands)."""
    with torch.cuda.device(_DEVICE):
        batch, labels, num, stats, stat_off, d_off = _label_image(img_bin)
        W = img_bin.shape[1]
        xs, ys, xf, yf = _seg.island_boxes(stats, W, margin=margin) if len(stats) else ([], [], [], [])
        groups = [np.array([k + 1]) for k in range(len(stats))]
        boxes = np.stack([xs, ys, xf, yf], axis=1).astype(np.int64) if len(stats) else np.zeros((0, 4), np.int64)
        crops = _seg.group_canvases(batch, labels, stat_off, d_off, [groups], [boxes])[0]
        islands = [(c, (int(top), int(left))) for c, (top, left) in crops]
        img_islands = batch.plane(labels, 0).cpu().numpy().copy()
    return islands, img_islands, num


def sort_islands(islands):
    """partition.py:90-98 (same np.argsort call; tie order is numpy's)."""
    order = np.argsort([isl[1][1] for isl in islands])
    return [islands[n] for n in order]


def group_intervals(intervals, width):
    """partition.py:248-318 (+ :321-358) through the native sd_group_intervals."""
    return _lib.group_intervals(intervals, width)


def group_islands(islands, target_shape):
    """partition.py:31-87 on a list of island crops (host; the batched device path
    is `segment.Segmenter.partition`, which never materialises per-island crops)."""
    islands = sort_islands(islands)
    intervals = [(pos[1], pos[1] + img.shape[1]) for img, pos in islands]
    out = []
    for grp in group_intervals(intervals, target_shape[1]):
        mem = [islands[k] for k in grp]
        left = np.min([p[1] for _, p in mem]); top = np.min([p[0] for _, p in mem])
        right = np.max([p[1] + im.shape[1] for im, p in mem]); bottom = np.max([p[0] + im.shape[0] for im, p in mem])
        canvas = np.zeros((bottom - top, right - left), np.uint8)
        for im, (r, c) in mem:
            canvas[r - top:r - top + im.shape[0], c - left:c - left + im.shape[1]] |= (im != 0).astype(np.uint8)
        out.append((canvas, (top, left)))
    return out


def get_pad_edges(n):
    """partition.py:241-245."""
    return (n // 2, n // 2) if n % 2 == 0 else (n // 2, n // 2 + 1)


def resize_and_pad_image(image, new_dims, margin=0, pad_value=0):
    """partition.py:101-140 (host cv2; the GPU crop generator is SURVEY.md 8(f) item 1)."""
    h, w = image.shape[:2]
    new_h, new_w = new_dims[0] - 2 * margin, new_dims[1] - 2 * margin
    scale = min(new_h / h, new_w / w)
    rs_w = int(np.min((np.rint(scale * w), new_w)))
    rs_h = int(np.min((np.rint(scale * h), new_h)))
    rs = cv2.resize(image, (rs_w, rs_h))
    ratio = (rs_w / w + rs_h / h) / 2
    ph = get_pad_edges(np.max((new_dims[0] - rs.shape[0], 0)))
    pw = get_pad_edges(np.max((new_dims[1] - rs.shape[1], 0)))
    pad = cv2.copyMakeBorder(rs, ph[0], ph[1], pw[0], pw[1], cv2.BORDER_CONSTANT, value=pad_value)
    return pad, ratio, ((pad.shape[1] - rs.shape[1]) / 2, (pad.shape[0] - rs.shape[0]) / 2)
