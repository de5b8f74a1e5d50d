"""Debug driver for an experimental conv kernel: one small forward, then the wait-error code."""
import os, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
from stroke_derenderer_b200 import _lib
from stroke_derenderer_b200.engine import UNetEngine
from stroke_derenderer_b200.weights import make_parity_weights
L = _lib.lib()
nt = int(os.environ.get("NT", "2"))
eng = UNetEngine(make_parity_weights(123), device=0, max_tiles=nt)
x = torch.rand((nt, 128, 384, 8), device="cuda").half(); x[..., 3:] = 0
m = torch.empty((nt, 128, 384), dtype=torch.uint8, device="cuda")
try:
    eng.enable_timing(True)
    eng.forward_into(x, m, 0.5)
    torch.cuda.synchronize()
    print("ok", [(n, round(t, 4)) for n, t in eng.layer_times()][:14])
except Exception as ex:
    print("FAILED:", str(ex)[:200])
print("wait error code:", L.sd_engine_wait_error(eng._h))

import ctypes as C
L.sd_engine_debug_word.restype = C.c_int; L.sd_engine_debug_word.argtypes = [C.c_void_p, C.c_int]
print("debug words:", [hex(L.sd_engine_debug_word(eng._h, i) & 0xffffffff) for i in range(10)])
