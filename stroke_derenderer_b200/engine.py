"""Host-side handle of the Attention-UNet engine (`sd_engine_*`, `sd_unet_forward`).

`UNetEngine` is what `BinarizationSession.init_onnx_inference()` returns in this
framework: it plays the role of the reference's `onnxruntime.InferenceSession`
(/root/reference/derenderer/evaluate_binarize.py:48-53) and keeps its one used
method, `.run(None, {"input": x}) -> [probabilities]` (:62, :100).  torch is used
only for device memory and streams.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .weights import fold_conv_bn, load_weights

TILE_H, TILE_W, CIN_PAD = 128, 384, 8


def _slot_sources():
    """slot name -> (conv prefix, bn prefix | None) in the upstream state_dict."""
    m = {}
    for n in range(1, 6):
        m[f"CONV{n}_0"] = (f"Conv{n}.conv.0", f"Conv{n}.conv.1")
        m[f"CONV{n}_1"] = (f"Conv{n}.conv.3", f"Conv{n}.conv.4")
    for n in range(5, 1, -1):
        m[f"UP{n}"] = (f"Up{n}.up.1", f"Up{n}.up.2")
        m[f"ATT{n}_G"] = (f"Att{n}.W_g.0", f"Att{n}.W_g.1")
        m[f"ATT{n}_X"] = (f"Att{n}.W_x.0", f"Att{n}.W_x.1")
        m[f"ATT{n}_PSI"] = (f"Att{n}.psi.0", f"Att{n}.psi.1")
        m[f"UPCONV{n}_0"] = (f"Up_conv{n}.conv.0", f"Up_conv{n}.conv.1")
        m[f"UPCONV{n}_1"] = (f"Up_conv{n}.conv.3", f"Up_conv{n}.conv.4")
    m["HEAD"] = ("Conv_1x1", None)
    return m


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class UNetEngine:
    """One engine per GPU; not thread-safe (include/sd_b200.h)."""

    def __init__(self, state, device: int = 0, max_tiles: int = 64, impl: int = 0):
        _lib.require_cuda()
        if isinstance(state, (str, bytes)) or hasattr(state, "__fspath__"):
            state = load_weights(str(state))
        self.device = torch.device("cuda", device)
        self.max_tiles = int(max_tiles)
        self.impl = impl
        self._h = C.c_void_p()
        L = _lib.lib()
        _lib.check(L.sd_engine_create(device, self.max_tiles, TILE_H, TILE_W, C.byref(self._h)), "sd_engine_create")
        srcs = _slot_sources()
        for i, name in enumerate(_lib.SLOTS):
            conv, bn = srcs[name]
            w, b = fold_conv_bn(state, conv, bn)
            w = np.ascontiguousarray(w, np.float32)
            b = np.ascontiguousarray(b, np.float32)
            _lib.check(L.sd_engine_set_conv(self._h, i, w.ctypes.data, b.ctypes.data, w.shape[0], w.shape[1], w.shape[2]),
                       f"sd_engine_set_conv({name})")
        with torch.cuda.device(self.device):
            _lib.check(L.sd_engine_finalize(self._h, impl), "sd_engine_finalize")
        self.head_bias = float(fold_conv_bn(state, "Conv_1x1", None)[1][0])

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            _lib.lib().sd_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_head_bias(self, bias: float):
        self.head_bias = float(bias)
        _lib.check(_lib.lib().sd_engine_set_head_bias(self._h, float(bias)), "sd_engine_set_head_bias")

    def enable_timing(self, on: bool = True):
        _lib.check(_lib.lib().sd_engine_enable_timing(self._h, int(on)))

    def layer_times(self):
        L = _lib.lib()
        n = C.c_int()
        buf = (C.c_float * 64)()
        _lib.check(L.sd_engine_layer_times(self._h, buf, 64, C.byref(n)))
        return [(L.sd_engine_layer_name(self._h, i).decode(), float(buf[i])) for i in range(n.value)]

    # ---- device-level forward -------------------------------------------------
    def forward(self, tiles: torch.Tensor, bin_thr: float = 0.5, want_prob32=False, want_prob16=False, want_mask=True):
        """tiles: (n, 128, 384, 8) fp16 NHWC on this GPU, n <= max_tiles.
        Returns dict with any of prob32 (n,128,384) f32, prob16 f16, mask u8 {0,255}."""
        assert tiles.is_cuda and tiles.dtype == _lib.torch_dtype() and tiles.is_contiguous()
        assert tuple(tiles.shape[1:]) == (TILE_H, TILE_W, CIN_PAD), tiles.shape
        n = tiles.shape[0]
        out = {}
        with torch.cuda.device(self.device):
            p32 = torch.empty((n, TILE_H, TILE_W), dtype=torch.float32, device=self.device) if want_prob32 else None
            p16 = torch.empty((n, TILE_H, TILE_W), dtype=torch.float16, device=self.device) if want_prob16 else None
            mk = torch.empty((n, TILE_H, TILE_W), dtype=torch.uint8, device=self.device) if want_mask else None
            _lib.check(_lib.lib().sd_unet_forward(
                self._h, tiles.data_ptr(), n, float(bin_thr),
                p32.data_ptr() if p32 is not None else None, p16.data_ptr() if p16 is not None else None,
                mk.data_ptr() if mk is not None else None, stream_ptr(self.device)), "sd_unet_forward")
        if p32 is not None: out["prob32"] = p32
        if p16 is not None: out["prob16"] = p16
        if mk is not None: out["mask"] = mk
        return out

    def forward_into(self, tiles: torch.Tensor, mask_out: torch.Tensor, bin_thr: float = 0.5):
        """Mask-only forward into a caller buffer slice (no allocation)."""
        n = tiles.shape[0]
        _lib.check(_lib.lib().sd_unet_forward(self._h, tiles.data_ptr(), n, float(bin_thr), None, None,
                                              mask_out.data_ptr(), stream_ptr(self.device)), "sd_unet_forward")

    def forward_lines(self, tiles: torch.Tensor, d_dst: torch.Tensor, bin_thr: float = 0.5):
        """Forward with the glue fused into the head (sd_unet_forward_lines): tile k ORs its thresholded,
        un-padded columns into the packed line planes named by entry k of `d_dst` (uint8 view of an
        sd_tile_dst table on this GPU).  The planes must have been zeroed on the same stream."""
        n = tiles.shape[0]
        assert d_dst.dtype == torch.uint8 and d_dst.numel() >= 16 * n
        _lib.check(_lib.lib().sd_unet_forward_lines(self._h, tiles.data_ptr(), n, float(bin_thr), d_dst.data_ptr(),
                                                    stream_ptr(self.device)), "sd_unet_forward_lines")

    def read_tap(self, name: str, n: int) -> torch.Tensor:
        """Intermediate activation of the last forward as (n, H, W, C) fp16."""
        L = _lib.lib()
        idx = _lib.TAPS.index(name)
        c, h, w = C.c_int(), C.c_int(), C.c_int()
        _lib.check(L.sd_unet_read_tap(self._h, idx, n, None, 0, C.byref(c), C.byref(h), C.byref(w), None), "sd_unet_read_tap")
        out = torch.empty((n, h.value, w.value, c.value), dtype=_lib.torch_dtype(), device=self.device)
        _lib.check(L.sd_unet_read_tap(self._h, idx, n, out.data_ptr(), out.numel() * 2, C.byref(c), C.byref(h), C.byref(w),
                                      stream_ptr(self.device)), "sd_unet_read_tap")
        return out

    # ---- the reference's model handle (B3) --------------------------------------
    @staticmethod
    def pack_input(x: torch.Tensor) -> torch.Tensor:
        """(B,3,H,W) f32 in [0,1] on GPU -> (B,H,W,8) fp16 NHWC, channels 3..7 zero."""
        b, c, h, w = x.shape
        t = torch.zeros((b, h, w, CIN_PAD), dtype=_lib.torch_dtype(), device=x.device)
        t[..., :c] = x.permute(0, 2, 3, 1).to(_lib.torch_dtype())
        return t

    def run(self, output_names, feeds):
        """onnxruntime-compatible call: `run(None, {"input": f32 (B,3,128,384)}) -> [f32 (B,1,128,384)]`.
        An empty batch returns an empty array (evaluate_binarize.py:93-100 feeds one when B % 8 == 0)."""
        x = np.asarray(feeds["input"])
        if x.ndim != 4 or x.shape[1] != 3 or tuple(x.shape[2:]) != (TILE_H, TILE_W):
            raise ValueError(f"input must be (B,3,{TILE_H},{TILE_W}), got {x.shape}")
        if x.shape[0] == 0:
            return [np.zeros((0, 1, TILE_H, TILE_W), np.float32)]
        outs = []
        with torch.cuda.device(self.device):
            for s in range(0, x.shape[0], self.max_tiles):
                xb = torch.from_numpy(np.ascontiguousarray(x[s:s + self.max_tiles], dtype=np.float32)).to(self.device)
                r = self.forward(self.pack_input(xb), want_prob32=True, want_mask=False)
                outs.append(r["prob32"].unsqueeze(1).cpu().numpy())
        return [np.concatenate(outs, 0)]
