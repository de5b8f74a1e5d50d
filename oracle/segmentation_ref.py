"""ORACLE (test infrastructure, never on the product path).

CPU restatement (numpy + cv2, same third-party calls as the reference) of the
reference's text-segmentation hot path.  Each function cites the reference
file:line it follows.  The restatement is pinned two ways (tests/test_oracle_*):
  * live, in the build container, against the unmodified reference functions
    imported from /root/reference (skipped where that tree is absent), and
  * against golden vectors under tests/golden/ that were produced by running
    the reference itself (tests/golden/make_golden.py, committed).
The reference has no tests or golden vectors of its own,
so for the UNet graph parity is "unpinned" (see oracle/attunet_torch.py); the
integer geometry here is pinned by the reference's own code.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import this file.
"""

from __future__ import annotations

import sys

import cv2
import numpy as np

HEIGHT, WIDTH, CHANNELS, OVERLAP, BIN_THR, MINIBATCH = 128, 384, 3, 64, 0.5, 8
IMAGENET_MEAN = [0.485, 0.456, 0.406]
IMAGENET_STD = [0.229, 0.224, 0.225]


# ----------------------------------------------------------------------------
# helper/split.py
# ----------------------------------------------------------------------------
def resize_to_height(img, height):
    """common.py:85-93 == helper/split.py:127-135: width=int(w*height/h),
    cv2.resize default interpolation (INTER_LINEAR)."""
    h, w = img.shape[:2]
    return cv2.resize(img, (int(w * (height / h)), height))


def pad_image(img, width, pad_value=0):
    """helper/split.py:42-54: right-pad with a constant to `width`, else crop."""
    extra = width - img.shape[1]
    if extra > 0:
        return cv2.copyMakeBorder(img, 0, 0, 0, extra, cv2.BORDER_CONSTANT, value=pad_value)
    return img[:, :width]


def tile_geometry(w, target_width=WIDTH, overlap=OVERLAP):
    """Closed form of helper/split.py:16-37 -> (starts, widths)."""
    if w < target_width:
        return [0], [w]
    n = w // (target_width - overlap) + 1
    wu = w // n
    starts = [i * wu for i in range(n)]
    widths = [min((i + 1) * wu + overlap, w) - i * wu for i in range(n)]
    return starts, widths


def split_image(img, target_width, overlap, pad_value=0):
    """helper/split.py:10-39."""
    starts, widths = tile_geometry(img.shape[1], target_width, overlap)
    tiles = [pad_image(img[:, s:s + wd], target_width, pad_value) for s, wd in zip(starts, widths)]
    return tiles, widths


def cut_and_stack(imgs_text, target_dim, overlap, pad_value=0):
    """helper/split.py:57-86 -> (stack (B,C,H,W) u8, stack_indices, stack_widths, img_widths)."""
    _, C, H, W = target_dim
    tiles, stack_indices, stack_widths, img_widths = [], [], [], []
    for img in imgs_text:
        rs = resize_to_height(img, H)
        cut, widths = split_image(rs, W, overlap, pad_value)
        stack_indices.append(list(range(len(tiles), len(tiles) + len(cut))))
        stack_widths.append(widths)
        img_widths.append(rs.shape[1])
        tiles.extend(cut)
    if C == 1:
        tiles = [t[:, :, None] for t in tiles]
    stack = np.stack([t.transpose(2, 0, 1) for t in tiles], axis=0)
    return stack, stack_indices, stack_widths, img_widths


def reconstruct_images(img_output, imgs_widths, stack_indices, stack_widths, overlap):
    """helper/split.py:89-124: un-pad, paste at stride width-overlap, max on overlaps."""
    _, C, H, _ = img_output.shape
    out = []
    for img_w, idxs, widths in zip(imgs_widths, stack_indices, stack_widths):
        canvas = np.zeros((H, img_w, C), np.uint8)
        x0 = 0
        for k, wd in zip(idxs, widths):
            piece = img_output[k][:, :, :wd].transpose(1, 2, 0)
            canvas[:, x0:x0 + wd, :] = np.maximum(canvas[:, x0:x0 + wd, :], piece)
            x0 += wd - overlap
        out.append(canvas)
    return out


# ----------------------------------------------------------------------------
# evaluate_binarize.py
# ----------------------------------------------------------------------------
class BinarizationSessionRef:
    """evaluate_binarize.py:26-150 (config semantics :30-45: JSON overrides kwargs)."""

    def __init__(self, configs_path=None, **params):
        if configs_path is not None:
            import json
            with open(configs_path) as f:
                params.update(json.load(f))
        self.height = params.get("height", HEIGHT)
        self.width = params.get("width", WIDTH)
        self.channels = params.get("channels", CHANNELS)
        self.overlap = params.get("overlap", OVERLAP)
        self.bin_thr = params.get("bin_thr", BIN_THR)
        self.minibatch = params.get("minibatch", MINIBATCH)

    def preprocess_images(self, images):
        """:67-82 (resize happens here and again inside cut_and_stack)."""
        rs = [resize_to_height(im, self.height) for im in images]
        return cut_and_stack(rs, (1, 3, self.height, self.width), self.overlap)

    def model_predict(self, img_stack, ort):
        """:85-115: B//mb+1 minibatches (the last may be empty), float64 /255
        then f32, strict `>` threshold, 255*u8."""
        B = img_stack.shape[0]
        outs = []
        for m in range(B // self.minibatch + 1):
            chunk = img_stack[m * self.minibatch:(m + 1) * self.minibatch]
            prob = ort.run(None, {"input": (chunk / 255.).astype(np.float32)})[0]
            bin_ = 255 * (prob > self.bin_thr).astype(np.uint8)
            if bin_.ndim == 3:
                bin_ = bin_[None]
            outs.append(bin_)
        return outs[0] if len(outs) == 1 else np.concatenate(outs, axis=0)

    def postprocess_stack(self, imgs_output, stack_indices, stack_widths, img_widths):
        """:118-127."""
        return reconstruct_images(imgs_output, img_widths, stack_indices, stack_widths, self.overlap)

    def binarize_images(self, images, ort):
        """:130-140."""
        stack, idx, widths, img_widths = self.preprocess_images(images)
        return self.postprocess_stack(self.model_predict(stack, ort), idx, widths, img_widths)

    def binarize_image(self, image, ort):
        """:143-150."""
        return self.binarize_images([image], ort)[0]


class TorchOrtSession:
    """Stand-in for onnxruntime.InferenceSession (evaluate_binarize.py:51-52,
    common.py:109-110): `.run(None, {"input": x})` -> [probabilities], executed
    by the torch-CPU fp32 oracle net.  NOT the real onnxruntime (absent here)."""

    def __init__(self, path_or_state, providers=None):
        from oracle.attunet_torch import build_oracle_net
        if isinstance(path_or_state, dict):
            state = path_or_state
        else:
            with np.load(path_or_state) as z:
                state = {k: z[k] for k in z.files}
        self.net = build_oracle_net(state)

    def run(self, output_names, feeds):
        from oracle.attunet_torch import oracle_unet_forward
        return [oracle_unet_forward(self.net, feeds["input"])]


def install_onnxruntime_shim():
    """Makes `import onnxruntime` work for the unmodified reference modules
    (SURVEY.md Appendix E).  A real onnxruntime, if ever present, wins."""
    try:
        import onnxruntime  # noqa: F401
        return False
    except ImportError:
        import types
        stub = types.ModuleType("onnxruntime")
        stub.InferenceSession = TorchOrtSession
        sys.modules["onnxruntime"] = stub
        return True


def post_glue_threshold(img_bin, bin_thr=BIN_THR):
    """main.py:108."""
    return img_bin[:, :, 0] > (255 * bin_thr)


# ----------------------------------------------------------------------------
# helper/partition.py (island half)
# ----------------------------------------------------------------------------
def get_binarized_islands(img_bin, margin=2):
    """helper/partition.py:9-28, including its per-island full-image scan."""
    num, labels, _, _ = cv2.connectedComponentsWithStats(img_bin)
    H, W = img_bin.shape[:2]
    islands = []
    for n in range(1, num):
        one = (labels == n).astype(np.uint8)
        x, y, w, h = cv2.boundingRect(one)
        xs, ys = max(x - margin, 0), max(y - margin, 0)
        xf, yf = min(x + w + margin + 1, W), min(y + h + margin + 1, H)
        islands.append((one[ys:yf, xs:xf], (ys, xs)))
    return islands, labels, num


def sort_islands(islands):
    """helper/partition.py:90-98 (np.argsort default kind on a Python list;
    tie order is numpy's)."""
    order = np.argsort([isl[1][1] for isl in islands])
    return [islands[i] for i in order]


def group_intervals(intervals, width):
    """helper/partition.py:248-318 with group_connections :321-345 and
    add_to_group :348-358 (recursive DFS replaced by an explicit stack that
    visits in the same order)."""
    N = len(intervals)
    adj = {n: [] for n in range(N)}
    contained = [False] * N
    for n, (a_o, b_o) in enumerate(intervals):
        if (b_o - a_o) <= width:
            continue
        for k, (a_i, b_i) in enumerate(intervals):
            if k == n:
                continue
            if a_i > b_o:
                break
            if a_o <= a_i and b_o >= b_i:
                adj[n].append(k)
                adj[k].append(n)
                contained[n] = contained[k] = True
    adj = {k: v for k, v in adj.items() if v}

    # group_connections + add_to_group: pre-order DFS from each unseen key in
    # insertion order; the start node itself is appended when first reached
    # from a neighbour (so it is not first in its own group).
    groups_long, done = [], set()
    for f in adj:
        if f in done:
            continue
        group, seen = [], set()
        stack = [iter(adj[f])]
        while stack:
            nxt = next(stack[-1], None)
            if nxt is None:
                stack.pop()
                continue
            if nxt not in seen:
                seen.add(nxt)
                group.append(nxt)
                stack.append(iter(adj[nxt]))
        done.update(group)
        done.add(f)
        groups_long.append(group)

    groups_short, cur, w, left = [], [], 0, 0
    for i, (a, b) in enumerate(intervals):
        if contained[i]:
            continue
        new_w = max(b - left, w)
        if new_w > width:
            groups_short.append(cur)
            cur, w, left = [i], b - a, a
        else:
            cur.append(i)
            w = new_w
    groups_short.append(cur)
    return [g for g in groups_long + groups_short if g]


def group_islands(islands, target_shape):
    """helper/partition.py:31-87 -> [(canvas u8 {0,1}, (top, left))]."""
    islands = sort_islands(islands)
    intervals = [(pos[1], pos[1] + img.shape[1]) for img, pos in islands]
    out = []
    for grp in group_intervals(intervals, target_shape[1]):
        members = [islands[k] for k in grp]
        left = np.min([p[1] for _, p in members])
        top = np.min([p[0] for _, p in members])
        right = np.max([p[1] + im.shape[1] for im, p in members])
        bottom = np.max([p[0] + im.shape[0] for im, p in members])
        canvas = np.zeros((bottom - top, right - left), np.uint8)
        for im, (r, c) in members:
            canvas[r - top:r - top + im.shape[0], c - left:c - left + im.shape[1]] += im.astype(np.uint8)
        out.append(((canvas > 0).astype(np.uint8), (top, left)))
    return out


def normalize_image(image):
    """common.py:96-102."""
    return cv2.normalize(image, None, 0, 255, norm_type=cv2.NORM_MINMAX)


def resize_linear_u8(src, dw, dh):
    """Restatement of cv2.resize(src, (dw, dh)) for one-channel uint8 with the default INTER_LINEAR, the call at
    helper/partition.py:120.  OpenCV (opencv-python, unpinned in the reference's setup.py:19; 4.13.0 in this
    image) is third-party; this follows its published 8-bit algorithm (imgproc/resize.cpp: fixed-point
    coefficients of INTER_RESIZE_COEF_BITS = 11 bits, HResizeLinear then VResizeLinear<uchar,int,short>) and is
    pinned by tests/test_oracle.py against cv2.resize itself on this image (bit-exact).
      * source position f = float((d + 0.5) * scale - 0.5), scale = 1 / (dst / src) in double; s = floor(f);
      * x: a source index outside [0, sw-1) is clamped AND its fraction zeroed; y: only the rows are clamped;
      * coefficients = round-half-even(float32 fraction * 2048) as int16;
      * vertical pass: (((b0 * (H0 >> 4)) >> 16) + ((b1 * (H1 >> 4)) >> 16) + 2) >> 2;
      * an exact 2x decimation in both directions takes the INTER_AREA fast path (2x2 mean, +2 >> 2)."""
    src = np.asarray(src, np.uint8)
    sh, sw = src.shape
    if sw == 2 * dw and sh == 2 * dh:
        s = src.astype(np.int32)
        return ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)

    def coeffs(dn, sn, clamp):
        scale = 1.0 / (dn / sn)
        ofs = np.zeros(dn, np.int64)
        a = np.zeros((dn, 2), np.int64)
        for d in range(dn):
            f = np.float32((d + 0.5) * scale - 0.5)
            s = int(np.floor(f))
            f = np.float32(f - np.float32(s))
            if clamp:
                if s < 0:
                    f, s = np.float32(0), 0
                if s >= sn - 1:
                    f, s = np.float32(0), sn - 1
            ofs[d] = s
            a[d, 0] = int(np.rint(np.float32((np.float32(1.0) - f) * np.float32(2048))))
            a[d, 1] = int(np.rint(np.float32(f * np.float32(2048))))
        return ofs, a

    xo, xa = coeffs(dw, sw, True)
    yo, ya = coeffs(dh, sh, False)
    S = src.astype(np.int64)
    H = S[:, xo] * xa[:, 0][None, :] + S[:, np.minimum(xo + 1, sw - 1)] * xa[:, 1][None, :]
    R0, R1 = H[np.clip(yo, 0, sh - 1), :], H[np.clip(yo + 1, 0, sh - 1), :]
    b0, b1 = ya[:, 0][:, None], ya[:, 1][:, None]
    out = (((b0 * (R0 >> 4)) >> 16) + ((b1 * (R1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def normalize_minmax_u8(img):
    """Restatement of cv2.normalize(img, None, 0, 255, NORM_MINMAX) for uint8 (common.py:100): scale and shift are
    computed in double, cast to float32, applied with ONE rounding (fused multiply-add) and rounded half-even;
    an all-equal image gives zeros.  Pinned against cv2 in tests/test_oracle.py."""
    img = np.asarray(img, np.uint8)
    mn, mx = int(img.min()), int(img.max())
    if mx == mn:
        return np.zeros_like(img)
    scale = 255.0 / (mx - mn)
    a, b = np.float32(scale), np.float32(0.0 - mn * scale)
    v = (img.astype(np.float64) * np.float64(a) + np.float64(b)).astype(np.float32)   # exact product + one rounding
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def get_pad_edges(n):
    """helper/partition.py:241-245."""
    return (n // 2, n // 2) if n % 2 == 0 else (n // 2, n // 2 + 1)


def resize_and_pad_image(image, new_dims, margin=0, pad_value=0):
    """helper/partition.py:101-140."""
    h, w = image.shape[:2]
    new_h, new_w = new_dims[0] - 2 * margin, new_dims[1] - 2 * margin
    scale = min(new_h / h, new_w / w)
    rs_w = int(np.min((np.rint(scale * w), new_w)))
    rs_h = int(np.min((np.rint(scale * h), new_h)))
    rs = cv2.resize(image, (rs_w, rs_h))
    ratio = (rs_w / w + rs_h / h) / 2
    ph = get_pad_edges(np.max((new_dims[0] - rs.shape[0], 0)))
    pw = get_pad_edges(np.max((new_dims[1] - rs.shape[1], 0)))
    padded = cv2.copyMakeBorder(rs, ph[0], ph[1], pw[0], pw[1], cv2.BORDER_CONSTANT, value=pad_value)
    return padded, ratio, ((padded.shape[1] - rs.shape[1]) / 2, (padded.shape[0] - rs.shape[0]) / 2)


def model_input_from_crop(img_u8, mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """evaluate_strokes.py:58-69 (_normalize_image)."""
    norm = normalize_image(img_u8.astype(np.uint8))
    return np.stack([(norm / 255. - mean[i]) / std[i] for i in range(3)], axis=0).astype(np.float32)


def get_partitions(img_bin, margin=2, img_size=224, mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """evaluate_strokes.py:186-224."""
    islands, _, _ = get_binarized_islands(img_bin.astype(np.uint8), margin=margin)
    h = img_bin.shape[0]
    parts = []
    for canvas, (top, left) in group_islands(islands, (h, h)):
        rs, ratio, (x2, y2) = resize_and_pad_image(normalize_image(canvas), (img_size, img_size),
                                                   margin=1, pad_value=0)
        parts.append({"image": rs, "image_input": model_input_from_crop(rs, mean, std),
                      "translate1": (left, top), "ratio": ratio, "translate2": (x2, y2)})
    return parts
