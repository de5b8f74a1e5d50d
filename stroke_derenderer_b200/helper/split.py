"""GPU-backed mirror of /root/reference/derenderer/helper/split.py: same function
names, arguments and return values; the cutting, padding, stacking and gluing run
as sm_100a kernels through the C ABI (sd_tile_extract_u8, sd_glue_u8).
"""

import numpy as np
import torch

from .. import segment as _seg
from ..common import resize_to_height  # noqa: F401  (split.py:127-135 re-exports it)

_DEVICE = 0


def set_device(device: int):
    global _DEVICE
    _DEVICE = int(device)


def _check_geometry(target_dim, overlap):
    _, C, H, W = target_dim
    if (H, W) != (_seg.TILE_H, _seg.TILE_W) or C != 3:
        raise ValueError(f"B200 path supports target_dim (B,3,{_seg.TILE_H},{_seg.TILE_W}); got {target_dim}")
    if not (0 <= overlap < W):
        raise ValueError(f"bad overlap {overlap}")


def split_image(img, target_width, overlap, pad_value=0):
    """split.py:10-39 for one image -> (list of padded tiles HWC, widths)."""
    stack, _, widths, _ = cut_and_stack([img], (1, img.shape[2] if img.ndim == 3 else 1, img.shape[0], target_width),
                                        overlap, pad_value)
    return [np.transpose(t, (1, 2, 0)) for t in stack], widths[0]


def cut_and_stack(imgs_text, target_dim, overlap, pad_value=0):
    """split.py:57-86 -> (img_stack (B,3,128,384) u8, stack_indices, stack_widths, img_widths)."""
    _check_geometry(target_dim, overlap)
    if pad_value != 0:
        raise ValueError("B200 path pads with 0 (the only value the reference uses)")
    _, C, H, W = target_dim
    rs = [resize_to_height(im, H) if im.shape[0] != H else im for im in imgs_text]
    dev = torch.device("cuda", _DEVICE)
    with torch.cuda.device(dev):
        batch = _seg.plan_batch([im.shape[1] for im in rs], dev, W, overlap)
        d_rgb = _seg.pack_lines_rgb(rs, batch).to(dev)
        stack = _seg.tile_extract_u8(batch, d_rgb).cpu().numpy()
    return stack, batch.stack_indices(), batch.stack_widths(), [im.shape[1] for im in rs]


def reconstruct_images(img_output, imgs_widths, stack_indices, stack_widths, overlap):
    """split.py:89-124 -> list of (128, W', 1) u8."""
    B, C, H, W = img_output.shape
    if C != 1 or (H, W) != (_seg.TILE_H, _seg.TILE_W):
        raise ValueError(f"B200 path glues (B,1,{_seg.TILE_H},{_seg.TILE_W}) outputs; got {img_output.shape}")
    dev = torch.device("cuda", _DEVICE)
    with torch.cuda.device(dev):
        batch = _seg.plan_batch(imgs_widths, dev, W, overlap)
        if batch.stack_indices() != [list(x) for x in stack_indices] or batch.n_tiles != B:
            raise ValueError("stack_indices do not describe a cut_and_stack of imgs_widths")
        tiles = torch.from_numpy(np.ascontiguousarray(img_output, dtype=np.uint8)).to(dev)
        planes = _seg.glue_u8(batch, tiles.view(B, H, W))
        return [batch.plane(planes, i).cpu().numpy()[:, :, None].copy() for i in range(batch.n_lines)]
