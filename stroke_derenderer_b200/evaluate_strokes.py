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


_host_pool = _seg.host_pool


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

    def get_partitions_batch(self, imgs_bin, lines_per_chunk: int = 64, keep_device: bool = False, lanes: int = 3):
        """Many lines in one pipelined device pass: per chunk of lines one H2D of the packed masks, CCL, stats,
        clustering, canvases and the 224x224 crops (sd_group_crops: cv2.normalize / cv2.resize / pad bit for bit),
        one D2H of the u8 crops straight into the page-locked arrays that are returned (no second host copy).
        A few host lanes (a thread + a CUDA stream each) take the chunks in turn, so packing / H2D / kernels / D2H of
        chunk k+1 run while the caller's thread builds the partition dicts of chunk k.
        -> per line the list of partition dicts of :213-219; `image_input` is built on first access
        (segment.LazyPartition), or stays on the GPU for the stroke-estimator front end (`keep_device=True` adds
        `self.last_device_crops`: per chunk the (g, 3, size, size) f32 tensor)."""
        if self.img_size % 2 or self.img_size > 256:
            raise ValueError(f"B200 crop kernel supports even image sizes up to 256, got {self.img_size} (no CPU fallback)")
        seg = self._segmenter()
        dev = seg.device
        lut = _seg.input_lut(self.mean, self.std)
        pool = _host_pool()
        fresh = _seg.FreshPinned()
        n_chunks = (len(imgs_bin) + lines_per_chunk - 1) // lines_per_chunk
        self.last_device_crops = [None] * n_chunks
        if not n_chunks:
            return []
        n_lanes = max(1, min(lanes, n_chunks, 4))
        streams = seg.lane_streams(n_lanes)
        with torch.cuda.device(dev):
            caller_stream = torch.cuda.current_stream(dev)
        from concurrent.futures import Future
        done = [Future() for _ in range(n_chunks)]

        def pack_one(args):
            m, ln, hn = args
            off, pitch, w = int(ln["px_off"]), int(ln["pitch"]), int(m.shape[1])
            view = hn[off:off + _seg.TILE_H * pitch].reshape(_seg.TILE_H, pitch)
            m = np.asarray(m)
            if m.dtype == np.bool_:
                view[:, :w] = m                       # img_bin.astype(np.uint8) of :193
            else:
                np.not_equal(m, 0, out=view[:, :w], casting="unsafe")
            view[:, w:] = 0                           # pad columns of the plane must be zero (CCL strips)

        def run_chunk(ci, lane):
            masks = imgs_bin[ci * lines_per_chunk:(ci + 1) * lines_per_chunk]
            key = ("lane", lane)                      # one staging set per lane: its chunks run one after the other
            batch = _seg.plan_batch([m.shape[1] for m in masks], dev)
            h = seg.staging.get_tensor((key, "masks"), batch.px_total)
            hn = h.numpy()
            list(pool.map(pack_one, [(m, ln, hn) for m, ln in zip(masks, batch.lines)]))   # numpy copies release the GIL
            planes = h.to(dev, non_blocking=True)
            res = seg.partition(batch, planes, canvases="device", key=key, zero_copy=True, crops=True, crops_to_host=True,
                                crop_lut=lut if keep_device else None, staging=fresh)
            if self.img_size != _seg.IMG_SIZE and len(res["groups"]):
                res["crops"] = _seg.group_crops(dev, res["canvas"], res["_keep"][0], res["groups"], size=self.img_size,
                                                lut=lut if keep_device else None)
                img = res["crops"]["image"]
                res["crops"]["image_host"] = _seg.copy_d2h(fresh.get(None, img.numel()), img, dev).reshape(tuple(img.shape))
            torch.cuda.current_stream(dev).synchronize()
            cr = res["crops"]
            if keep_device and cr is not None and cr["image_input"] is not None:
                cr["image_input"].record_stream(caller_stream)     # allocated on the lane's stream, read on the caller's
                self.last_device_crops[ci] = cr["image_input"]
            return (cr["image_host"] if cr is not None and "image_host" in cr else None, res["groups"], cr,
                    res["line_group_start"], batch.n_lines)

        def run_lane(lane):
            try:
                with torch.cuda.device(dev), torch.cuda.stream(streams[lane]):
                    for ci in range(lane, n_chunks, n_lanes):
                        done[ci].set_result(run_chunk(ci, lane))
            except BaseException as e:                # the caller's thread re-raises it from the first unfinished chunk
                for f in done:
                    if not f.done():
                        f.set_exception(e)

        with torch.cuda.device(dev):
            cur = torch.cuda.current_stream(dev)
            for st in streams:
                st.wait_stream(cur)
            lanes = [_seg.lane_pool().submit(run_lane, k) for k in range(n_lanes)]
            out = []
            try:
                for f in done:
                    imgs_h, groups, cr, lgs, n_lines = f.result()
                    out.extend(_seg.build_partitions(lut, imgs_h, groups, cr, lgs, n_lines))
            finally:
                for ln in lanes:
                    ln.result()
                for st in streams:
                    cur.wait_stream(st)
        return out

    # ---- stroke-estimator front end (evaluate_strokes.py:150-160, 250-262) --------------------
    def load_orts(self, filepaths):
        """:150-160.  The stroke-estimator graphs (encoder, projection, decoder_*) live on the same Drive link as the
        binarizer and, unlike the binarizer, not even their topology is named in the reference, so this framework
        cannot execute them: every value of `filepaths` must already be a model HANDLE, an object with
        `.run(output_names, {"input": ndarray}) -> [ndarray]` (the onnxruntime signature the reference uses) and
        optionally `.run_device(output_names, {"input": cuda tensor}) -> [cuda tensor]` to keep the batch in HBM.
        A path raises: there is no CPU fallback and no ONNX executor here."""
        orts = {}
        for k, v in filepaths.items():
            if not hasattr(v, "run"):
                raise NotImplementedError(f"'{k}': {v!r} is not a model handle; the stroke-estimator graphs are not part of the "
                                          "reference tree and cannot be executed by this framework (SURVEY.md 2)")
            orts[k] = v
        return orts

    @staticmethod
    def _run_handle(handle, out_names, x):
        """One model call with the batch kept on the GPU when the handle can take it."""
        if hasattr(handle, "run_device"):
            return handle.run_device(out_names, {"input": x})[0]
        y = handle.run(out_names, {"input": x.cpu().numpy() if isinstance(x, torch.Tensor) else x})[0]
        return torch.from_numpy(np.ascontiguousarray(y)).to(x.device) if isinstance(x, torch.Tensor) else y

    def _encode_postprocess(self, enc):
        """:72-91 on the GPU (sd_encode_postprocess): (B, C, E/2, E/2) -> (B, E*E, C) f32, values repeated on a 2x2
        grid.  numpy in -> numpy out, cuda tensor in -> cuda tensor out."""
        from . import _lib
        is_np = not isinstance(enc, torch.Tensor)
        dev = torch.device("cuda", self.device)
        t = torch.from_numpy(np.ascontiguousarray(enc, dtype=np.float32)).to(dev) if is_np else enc.to(torch.float32).contiguous()
        B, C, h, w = t.shape
        E = self.enc_image_size
        if (2 * h, 2 * w) != (E, E):
            raise ValueError(f"encoder output {h}x{w} does not fill the {E}x{E} grid of _encode_postprocess")
        out = torch.empty((B, E * E, C), dtype=torch.float32, device=t.device)
        with torch.cuda.device(t.device):
            _lib.check(_lib.lib().sd_encode_postprocess(t.data_ptr(), B, C, h, w, out.data_ptr(),
                                                        torch.cuda.current_stream(t.device).cuda_stream), "sd_encode_postprocess")
        return out.cpu().numpy() if is_np else out

    def encode_partitions_batch(self, imgs_bin, orts, max_batch: int = 512):
        """The front half of `estimate_strokes` (:250-262) for MANY lines at once: partitions of every line (device
        crops, never copied to the host as f32) -> encoder in batches of up to `max_batch` crops that cross line
        boundaries -> _encode_postprocess -> optional projection.  The reference runs this per line image with
        whatever batch that line happens to have (:171-181).
        -> (partitions per line, enc per line: cuda tensor (n_parts, P, E) f32)."""
        parts = self.get_partitions_batch(imgs_bin, keep_device=True)
        crops = [c for c in self.last_device_crops if c is not None and c.shape[0]]
        encs = []
        if crops:
            allc = torch.cat(crops, 0) if len(crops) > 1 else crops[0]
            for s0 in range(0, allc.shape[0], max_batch):
                enc = self._run_handle(orts["encoder"], ["output"], allc[s0:s0 + max_batch])
                enc = self._encode_postprocess(enc)
                if "projection" in orts:
                    enc = self._run_handle(orts["projection"], ["output"], enc)
                encs.append(enc)
        enc_all = torch.cat(encs, 0) if encs else None
        out, g = [], 0
        for pl in parts:
            out.append(enc_all[g:g + len(pl)] if enc_all is not None else torch.zeros((0, 0, 0)))
            g += len(pl)
        return parts, out

    def process_image(self, img_bin, orts, max_length=None):
        raise NotImplementedError("the LSTM-attention stroke decoder (evaluate_strokes.py:262-303) is a second model outside the "
                                  "B200 segmentation path (SURVEY.md 2); its front end is encode_partitions_batch")
