"""Segmentation half of /root/reference/derenderer/evaluate_strokes.py:
`StrokeEstimationSession.get_partitions` (:186-224) with labelling, island boxes
and group canvases on the B200.  The stroke-estimator networks (encoder / LSTM
decoder, :250-303) are a second model outside this hot path.
"""

import numpy as np
import torch

from . import segment as _seg
from .common import load_json, normalize_image
from .helper.partition import resize_and_pad_image

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

    def partitions_from_canvases(self, canvases):
        """:202-222 for one line's [(canvas, (top, left))]."""
        parts = []
        for img, (y, x) in canvases:
            img_rs, ratio, (x2, y2) = resize_and_pad_image(normalize_image(img), self.tgt_shape, margin=1, pad_value=0)
            parts.append({"image": img_rs, "image_input": self._normalize_image(img_rs),
                          "translate1": (x, y), "ratio": ratio, "translate2": (x2, y2)})
        return parts

    def get_partitions(self, img_bin):
        """:186-224 for one (128, W) binary image."""
        return self.get_partitions_batch([img_bin])[0]

    def get_partitions_batch(self, imgs_bin):
        """Many lines in one device pass (CCL, stats, canvases batched)."""
        dev = torch.device("cuda", self.device)
        with torch.cuda.device(dev):
            batch = _seg.plan_batch([m.shape[1] for m in imgs_bin], dev)
            host = np.zeros(batch.px_total, np.uint8)
            for m, ln in zip(imgs_bin, batch.lines):
                off, pitch = int(ln["px_off"]), int(ln["pitch"])
                host[off:off + _seg.TILE_H * pitch].reshape(_seg.TILE_H, pitch)[:, :m.shape[1]] = np.asarray(m) != 0
            planes = torch.from_numpy(host).to(dev)
            if self.img_size % 2 or self.img_size > 256:
                res = _seg.Segmenter(None, margin=self.margin, device=dev).partition(batch, planes)
                return [self.partitions_from_canvases(c) for c in res.canvases]
            # crops on the device too (sd_group_crops): cv2.normalize / cv2.resize / pad / mean-std bit for bit
            seg = _seg.Segmenter(None, margin=self.margin, device=dev)
            res = seg.partition(batch, planes, canvases="device", crops=True, crop_lut=_seg.input_lut(self.mean, self.std))
            if self.img_size != _seg.IMG_SIZE and len(res["groups"]):
                res["crops"] = _seg.group_crops(dev, res["canvas"], res["_keep"][0], res["groups"], size=self.img_size,
                                                lut=_seg.input_lut(self.mean, self.std))
            return [res.line_partitions(l) for l in range(batch.n_lines)]

    def load_orts(self, filepaths):
        raise NotImplementedError("stroke-estimator graphs are outside the B200 segmentation path (SURVEY.md 2)")

    def process_image(self, img_bin, orts, max_length=None):
        raise NotImplementedError("stroke estimation is outside the B200 segmentation path (SURVEY.md 2)")
