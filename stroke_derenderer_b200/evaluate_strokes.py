"""Segmentation half of /root/reference/derenderer/evaluate_strokes.py:
`StrokeEstimationSession.get_partitions` (:186-224) with labelling, island boxes
and group canvases on the B200.  The stroke-estimator networks (encoder / LSTM
decoder, :250-303) are a second model outside this hot path.
"""

import numpy as np
import torch

from . import segment as _seg
from .common import load_json, normalize_image

# evaluate_strokes.py:24-31
IMG_SIZE = 224
MARGIN = 2
MAX_LENGTH = 384
MEAN = [0.485, 0.456, 0.406]
STD = [0.229, 0.224, 0.225]
PAD, BOS, EOS = 0, 1, 2


class StrokeEstimationSession:
    def __init__(self, configs_path=None, **params):
        # :35-50
        if configs_path is not None:
            params.update(load_json(configs_path))
        self.max_length = params.get("max_length", MAX_LENGTH)
        self.img_size = params.get("image_size", IMG_SIZE)
        self.margin = params.get("margin", MARGIN)
        self.mean = params.get("mean", MEAN)
        self.std = params.get("std", STD)
        self.enc_image_size = params.get("encode_image_size", 14)
        self.device = params.get("device", 0)

    @property
    def tgt_shape(self):
        return (self.img_size, self.img_size)

    def _normalize_image(self, img_bin):
        """:58-69."""
        img_norm = normalize_image(img_bin.astype(np.uint8))
        return np.stack([(img_norm / 255. - self.mean[i]) / self.std[i] for i in range(3)], axis=0).astype(np.float32)

    def get_partitions(self, img_bin):
        """:186-224 for one (128, W) binary image."""
        return self.get_partitions_batch([img_bin])[0]

    def _segmenter(self):
        dev = torch.device("cuda", self.device)
        seg = getattr(self, "_seg_obj", None)
        if seg is None or seg.margin != self.margin:
            seg = self._seg_obj = _seg.Segmenter(None, margin=self.margin, device=dev)
        return seg

    def get_partitions_batch(self, imgs_bin, lines_per_chunk: int = 64, keep_device: bool = False):
        """Many lines in one pipelined device pass: per chunk of lines one H2D of the packed masks, CCL, stats,
        clustering, canvases and the 224x224 crops (sd_group_crops: cv2.normalize / cv2.resize / pad bit for bit),
        one D2H of the u8 crops.  -> per line the list of partition dicts of :213-219; `image_input` is built on
        first access (segment.LazyPartition), or stays on the GPU for the stroke-estimator front end
        (`keep_device=True` adds `self.last_device_crops`: per chunk the (g, 3, size, size) f32 tensor)."""
        if self.img_size % 2 or self.img_size > 256:
            raise ValueError(f"B200 crop kernel supports even image sizes up to 256, got {self.img_size} (no CPU fallback)")
        seg = self._segmenter()
        dev = seg.device
        lut = _seg.input_lut(self.mean, self.std)
        out = []
        self.last_device_crops = []
        with torch.cuda.device(dev):
            for c0 in range(0, len(imgs_bin), lines_per_chunk):
                masks = imgs_bin[c0:c0 + lines_per_chunk]
                key = ("chunk", (c0 // lines_per_chunk) & 1)           # two staging sets: pack k+1 while k is in flight
                batch = _seg.plan_batch([m.shape[1] for m in masks], dev)
                h = seg.staging.get_tensor((key, "masks"), batch.px_total)
                hn = h.numpy()
                for m, ln in zip(masks, batch.lines):
                    off, pitch, w = int(ln["px_off"]), int(ln["pitch"]), int(m.shape[1])
                    view = hn[off:off + _seg.TILE_H * pitch].reshape(_seg.TILE_H, pitch)
                    m = np.asarray(m)
                    if m.dtype == np.bool_:
                        view[:, :w] = m                   # img_bin.astype(np.uint8) of :193
                    else:
                        np.not_equal(m, 0, out=view[:, :w], casting="unsafe")
                    view[:, w:] = 0                       # pad columns of the plane must be zero (CCL strips)
                planes = h.to(dev, non_blocking=True)
                res = seg.partition(batch, planes, canvases="device", key=key, crops=True, crops_to_host=True,
                                    crop_lut=lut if keep_device else None)
                if self.img_size != _seg.IMG_SIZE and len(res["groups"]):
                    res["crops"] = _seg.group_crops(dev, res["canvas"], res["_keep"][0], res["groups"], size=self.img_size,
                                                    lut=lut if keep_device else None)
                    img = res["crops"]["image"]
                    res["crops"]["image_host"] = _seg.copy_d2h(seg.staging.get((key, "crops"), img.numel()), img, dev).reshape(tuple(img.shape))
                torch.cuda.current_stream(dev).synchronize()
                cr, groups, lgs = res["crops"], res["groups"], res["line_group_start"]
                if keep_device:
                    self.last_device_crops.append(cr["image_input"] if cr is not None else None)
                imgs_h = cr["image_host"].copy() if cr is not None and "image_host" in cr else None
                if cr is not None and len(groups):
                    left, top = groups[:, 1], groups[:, 2]
                    ratio, t2x, t2y = cr["ratio"].tolist(), cr["translate2"][:, 0].tolist(), cr["translate2"][:, 1].tolist()
                LP = _seg.LazyPartition
                for k in range(batch.n_lines):
                    out.append([LP(lut, image=imgs_h[g], translate1=(left[g], top[g]), ratio=ratio[g], translate2=(t2x[g], t2y[g]))
                                for g in range(int(lgs[k]), int(lgs[k + 1]))])
        return out

    def load_orts(self, filepaths):
        raise NotImplementedError("stroke-estimator graphs are outside the B200 segmentation path (SURVEY.md 2)")

    def process_image(self, img_bin, orts, max_length=None):
        raise NotImplementedError("stroke estimation is outside the B200 segmentation path (SURVEY.md 2)")
