"""Drop-in for /root/reference/derenderer/evaluate_binarize.py: the same
`BinarizationSession` (constructor, attributes, method names, argument and return
conventions) with every stage on the B200.  `init_onnx_inference` returns a
`UNetEngine` instead of an onnxruntime session; callers pass it back as `ort`.
"""

import numpy as np
import torch

from . import segment as _seg
from .common import load_json, resize_to_height
from .engine import UNetEngine
from .helper.split import cut_and_stack, reconstruct_images

# evaluate_binarize.py:19-24
HEIGHT = 128
WIDTH = 128 * 3
CHANNELS = 3
OVERLAP = 128 // 2
BIN_THR = 0.5
MINIBATCH = 8


class BinarizationSession:
    def __init__(self, configs_path=None, **params):
        # :30-45 — the JSON file overrides keyword arguments
        if configs_path is not None:
            params.update(load_json(configs_path))
        self.height = params.get("height", HEIGHT)
        self.width = params.get("width", WIDTH)
        self.channels = params.get("channels", CHANNELS)
        self.overlap = params.get("overlap", OVERLAP)
        self.bin_thr = params.get("bin_thr", BIN_THR)
        self.minibatch = params.get("minibatch", MINIBATCH)
        self.device = params.get("device", 0)
        self.max_tiles = params.get("max_tiles", 64)
        self.lines_per_chunk = params.get("lines_per_chunk", 32)

    def init_onnx_inference(self, onnxpath):
        """:48-53.  `onnxpath`: an exported `binarizer.onnx` (initializers read by `onnx_reader`, no
        onnx/onnxruntime needed), a `.npz` state dict with the upstream parameter names, or a dict of arrays."""
        return UNetEngine(onnxpath, device=self.device, max_tiles=self.max_tiles)

    def ort_predict(self, input_numpy, ort):
        """:56-64."""
        return ort.run(None, {"input": input_numpy})[0]

    def preprocess_images(self, images):
        """:67-82."""
        images_rs = [resize_to_height(im, self.height) if im.shape[0] != self.height else im for im in images]
        return cut_and_stack(images_rs, (1, 3, self.height, self.width), self.overlap)

    def model_predict(self, img_stack, ort):
        """:85-115 -> (B,1,128,384) u8 in {0,255}.  The reference's loop over
        B//minibatch+1 host minibatches (the last possibly empty) collapses into
        engine-sized device batches; `minibatch` is kept as an attribute only."""
        B = img_stack.shape[0]
        dev = ort.device
        outs = []
        with torch.cuda.device(dev):
            for s in range(0, B, ort.max_tiles):
                x = torch.from_numpy(np.ascontiguousarray(img_stack[s:s + ort.max_tiles])).to(dev)
                # (img / 255.).astype(float32) (:99), then the engine's fp16 NHWC input layout
                t = ort.pack_input(x.to(torch.float32) / 255.0)
                outs.append(ort.forward(t, bin_thr=self.bin_thr)["mask"].unsqueeze(1).cpu().numpy())
        if not outs:
            return np.zeros((0, 1, self.height, self.width), np.uint8)
        return outs[0] if len(outs) == 1 else np.concatenate(outs, axis=0)

    def postprocess_stack(self, imgs_output, stack_indices, stack_widths, img_widths):
        """:118-127."""
        return reconstruct_images(imgs_output, img_widths, stack_indices, stack_widths, self.overlap)

    def binarize_images(self, images, ort):
        """:130-140 -> list of (128, W', 1) u8 {0,255}.  Fused device path: nothing
        but the input lines and the glued masks crosses PCIe."""
        if (self.height, self.width, self.overlap) != (HEIGHT, WIDTH, OVERLAP):
            raise ValueError("B200 path supports the default 128/384/64 geometry")
        if not len(images):
            return []
        seg = self._segmenter(ort)
        # chunks of lines flow through pack -> H2D -> cut -> UNet (glue fused) -> ONE D2H of the packed planes per
        # chunk; resize_to_height (:76) happens on the device, bit-exact with cv2
        from .pipeline import LineSegmentationJob
        job = LineSegmentationJob(ort, images, bin_thr=self.bin_thr, lines_per_chunk=self.lines_per_chunk, crops=False,
                                  prepack=False, seg=seg)
        return job.binarize_step()

    def _segmenter(self, ort):
        """One Segmenter (streams, staging buffers) per engine, kept across calls."""
        return _seg.Segmenter.for_engine(ort, self.bin_thr)

    def binarize_image(self, image, ort):
        """:143-150."""
        return self.binarize_images([image], ort)[0]
