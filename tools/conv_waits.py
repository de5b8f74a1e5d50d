"""Debug: where do the warp roles of the tcgen05 conv kernels wait?  Needs a library built with
SD_EXTRA_NVCC_FLAGS=-DSD_CONV_STATS.  Runs one warmed UNet pass layer by layer is not possible through the C ABI,
so it reports the totals of a whole pass and, with SD_ONLY=<substring>, nothing else changes: use the per-layer
event times next to it."""
import ctypes as C, json, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
from stroke_derenderer_b200 import _lib
from stroke_derenderer_b200.engine import UNetEngine
from stroke_derenderer_b200.weights import make_parity_weights

L = _lib.lib(); _lib.require_cuda()
nt = 128
state = make_parity_weights(123)
eng = UNetEngine(state, device=0, max_tiles=nt)
tiles = torch.rand((nt, 128, 384, 8), device="cuda").half(); tiles[..., 3:] = 0
masks = torch.empty((nt, 128, 384), dtype=torch.uint8, device="cuda")
for _ in range(2): eng.forward_into(tiles, masks, 0.5)
torch.cuda.synchronize()
buf = (C.c_ulonglong * 8)()
L.sd_debug_wait_cycles(buf, 1)
eng.enable_timing(True)
eng.forward_into(tiles, masks, 0.5)
torch.cuda.synchronize()
L.sd_debug_wait_cycles(buf, 1)
print(json.dumps({"layers_ms": {n: round(t, 4) for n, t in eng.layer_times()}, "wait_cycles": list(buf)}))
