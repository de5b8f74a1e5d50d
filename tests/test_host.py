"""CPU-side checks: the C-ABI library builds/loads and exports every symbol the
header declares, host planners match the oracle, and the product path fails
loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from oracle import segmentation_ref as O
from stroke_derenderer_b200 import _lib

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol():
    hdr = (ROOT / "include" / "sd_b200.h").read_text()
    declared = set(re.findall(r"\b(sd_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 24
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), f"libsd_b200.so does not export {name}"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert L.sd_version() >= 100


def test_line_struct_layout_matches_header():
    hdr = (ROOT / "include" / "sd_b200.h").read_text()
    body = hdr[hdr.index("typedef struct sd_line {"):hdr.index("} sd_line;")]
    fields = re.findall(r"\b(int64_t|int32_t)\s+(\w+);", body)
    assert [f for _, f in fields] == list(_lib.LINE_DTYPE.names)
    assert sum(8 if t == "int64_t" else 4 for t, _ in fields) == _lib.LINE_DTYPE.itemsize


def test_resize_job_layout_and_plan():
    """struct sd_resize_job == its numpy mirror, and segment.ResizePlan builds the table the kernel's contract asks
    for: destination widths and slots of resize_to_height (common.py:89-91), every source image followed by at
    least 8 readable bytes, 128-px lines left to the host pack."""
    import torch
    from stroke_derenderer_b200 import segment as S
    hdr = (ROOT / "include" / "sd_b200.h").read_text()
    body = hdr[hdr.index("typedef struct sd_resize_job {"):hdr.index("} sd_resize_job;")]
    names = []
    for t, decl in re.findall(r"\b(int64_t|int32_t)\s+([\w\s,]+);", body):
        names += [(t, n.strip()) for n in decl.split(",")]
    assert [n for _, n in names] == list(_lib.RESIZE_DTYPE.names)
    assert sum(8 if t == "int64_t" else 4 for t, _ in names) == _lib.RESIZE_DTYPE.itemsize == 32
    rng = np.random.default_rng(0)
    shapes = [(200, 1000), (128, 700), (64, 300), (256, 1024), (37, 401)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes]
    widths = [S.resized_width(im) for im in imgs]
    assert widths == [O.resize_to_height(im, 128).shape[1] for im in imgs]
    lines, plan = _lib.plan_lines(widths)
    batch = S.LineBatch(torch.device("cpu"), lines, plan, torch.zeros(1, dtype=torch.uint8), widths)
    rp = S.ResizePlan(imgs, batch, pinned=False)
    jobs = rp.d_jobs.numpy().view(_lib.RESIZE_DTYPE)
    assert rp.n == 4 and rp.max_dst_w == max(w for w, im in zip(widths, imgs) if im.shape[0] != 128)
    todo = [i for i, im in enumerate(imgs) if im.shape[0] != 128]
    end = 0
    for jb, i in zip(jobs, todo):
        im = imgs[i]
        assert (int(jb["src_h"]), int(jb["src_w"]), int(jb["dst_w"])) == (im.shape[0], im.shape[1], widths[i])
        assert int(jb["dst_off"]) == int(lines[i]["img_off"]) and int(jb["src_off"]) % 16 == 0 and int(jb["src_off"]) >= end
        a = int(jb["src_off"])
        assert np.array_equal(rp.h_src.numpy()[a:a + im.size], im.reshape(-1))
        end = a + im.size + 8
    assert rp.h_src.numel() >= end
    packed = S.pack_lines_rgb(imgs, batch, pinned=False).numpy()
    off = int(lines[1]["img_off"])
    assert np.array_equal(packed[off:off + imgs[1].size], imgs[1].reshape(-1))


def test_plan_lines_matches_oracle_geometry():
    widths = [1, 2, 100, 383, 384, 385, 639, 640, 1000, 1536, 3072, 6144, 16384, 20480, 21000, 32768]
    lines, plan = _lib.plan_lines(widths)
    t = 0
    for ln, W in zip(lines, widths):
        starts, ws = O.tile_geometry(W)
        assert ln["n_tiles"] == len(ws) and ln["first_tile"] == t
        if len(ws) > 1:
            assert ln["wu"] == starts[1]
        assert ln["pitch"] % 128 == 0 and ln["pitch"] >= W and ln["bw"] * 2 == ln["pitch"]
        assert ln["px_off"] % 16 == 0 and ln["blk_off"] % 4096 == 0 and ln["img_off"] % 16 == 0
        t += len(ws)
    assert plan.n_tiles == t and plan.n_lines == len(widths)
    assert plan.px_total == int(sum(128 * ln["pitch"] for ln in lines))


def test_plan_lines_rejects_bad_input():
    with pytest.raises(_lib.SdError):
        _lib.plan_lines([100, 0])
    with pytest.raises(_lib.SdError):
        _lib.plan_lines([100], tile_w=64, overlap=64)


def test_native_group_intervals_golden(golden):
    for case in golden["group_intervals"]:
        assert _lib.group_intervals(case["intervals"], 128) == case["groups"]


def test_native_group_intervals_random_vs_oracle():
    rng = np.random.default_rng(5)
    for _ in range(500):
        n = int(rng.integers(0, 80))
        a = np.sort(rng.integers(0, 2000, n))
        w = np.where(rng.random(n) < 0.2, rng.integers(129, 700, n), rng.integers(1, 100, n))
        iv = [(int(x), int(x + y)) for x, y in zip(a, w)]
        assert _lib.group_intervals(iv, 128) == O.group_intervals(iv, 128)


def test_group_line_matches_oracle_group_islands(golden, golden_arrays):
    """stats closed form == get_binarized_islands + group_islands, from cv2 stats."""
    import cv2
    from stroke_derenderer_b200.segment import group_line
    for name in ["line300", "line1000", "line3072", "dense2048", "long_island"]:
        shp = golden["islands"][name]["shape"]
        m = np.unpackbits(golden_arrays[f"{name}_mask"])[:shp[0] * shp[1]].reshape(shp)
        n, labels, stats, _ = cv2.connectedComponentsWithStats(m)
        groups, boxes = group_line(stats[1:], shp[1])
        want = golden["islands"][name]["groups"]
        assert len(groups) == len(want)
        for members, (l, t, r, b), g in zip(groups, boxes, want):
            assert [int(t), int(l)] == g["pos"] and [int(b - t), int(r - l)] == g["shape"]
            canvas = np.isin(labels[t:b, l:r], members).astype(np.uint8)
            import hashlib
            assert hashlib.sha256(canvas.tobytes()).hexdigest() == g["sha"]


def test_no_cpu_fallback():
    """Without a CUDA device the compute entry points must fail, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    L = _lib.lib()
    assert L.sd_cuda_available() == 0
    h = C.c_void_p()
    assert L.sd_engine_create(0, 8, 128, 384, C.byref(h)) == -2      # SD_ECUDA
    assert b"no CUDA device" in L.sd_last_error()
    from stroke_derenderer_b200.engine import UNetEngine
    with pytest.raises(_lib.SdError):
        UNetEngine({}, device=0)


def test_product_never_imports_oracle():
    pkg = ROOT / "stroke_derenderer_b200"
    for p in list(pkg.rglob("*.py")) + [ROOT / "main.py"]:
        if p.exists():
            assert not re.search(r"^\s*(from|import)\s+oracle\b", p.read_text(), re.M), p


def test_session_config_semantics(tmp_path):
    """evaluate_binarize.py:30-45: defaults, kwargs, and JSON overriding kwargs."""
    import json
    from stroke_derenderer_b200.evaluate_binarize import BinarizationSession
    bs = BinarizationSession()
    assert (bs.height, bs.width, bs.channels, bs.overlap, bs.bin_thr, bs.minibatch) == (128, 384, 3, 64, 0.5, 8)
    cfg = tmp_path / "c.json"
    cfg.write_text(json.dumps({"bin_thr": 0.3}))
    bs = BinarizationSession(configs_path=str(cfg), bin_thr=0.9, minibatch=4)
    assert bs.bin_thr == 0.3 and bs.minibatch == 4


from onnx_writer import write_onnx as _write_onnx   # minimal ONNX (protobuf wire format) writer, tests/onnx_writer.py


def test_onnx_reader_recovers_folded_weights(tmp_path):
    """An exported binarizer.onnx (BN folded by the exporter, or left as BatchNormalization nodes) yields the
    same folded conv weights as the state dict it was made from."""
    from stroke_derenderer_b200.onnx_reader import load_onnx_state
    from stroke_derenderer_b200.weights import conv_bn_slots, fold_conv_bn, load_weights, make_parity_weights
    state = make_parity_weights(5)
    for folded in (True, False):
        p = tmp_path / f"binarizer_{int(folded)}.onnx"
        _write_onnx(p, state, folded)
        got = load_weights(str(p))
        assert set(got) == {f"{c}.{k}" for c, *_ in conv_bn_slots() for k in ("weight", "bias")}
        for conv, bn, *_ in conv_bn_slots():
            w, b = fold_conv_bn(state, conv, bn)
            gw, gb = fold_conv_bn(got, conv, bn)                 # no BN entries -> returned as is
            assert np.array_equal(gw, w) if folded else np.allclose(gw, w, rtol=1e-6, atol=1e-7), conv
            assert np.array_equal(gb, b) if folded else np.allclose(gb, b, rtol=1e-6, atol=1e-7), conv
    bad = tmp_path / "bad.onnx"
    bad.write_bytes(b"\x08\x08")
    with pytest.raises(ValueError):
        load_onnx_state(str(bad))
    # graphs with the right conv shapes but another wiring load cleanly nowhere: the engine hard-codes the data flow
    for kw, what in (({"final_sigmoid": False}, "Sigmoid"), ({"resize_mode": "linear"}, "nearest"), ({"swap_concat": True}, "gated skip")):
        p = tmp_path / "wrong.onnx"
        _write_onnx(p, state, True, **kw)
        with pytest.raises(ValueError, match=what):
            load_onnx_state(str(p))
        assert len(load_onnx_state(str(p), check=False)) == 2 * len(conv_bn_slots())      # the weights themselves still parse


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) needs no GPU and prints exactly
    one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "tiles/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_fastdiv_identity():
    """csrc/common.cuh FastDiv: q = (x * (2^40 // d + 1)) >> 40 is floor(x / d) for every x, d < 2^20 (the conv kernels
    decompose work-item and tile indices with it); checked on the edges and on random pairs."""
    rng = np.random.default_rng(0)
    ds = np.concatenate([np.arange(1, 300), rng.integers(1, 1 << 20, 4000), [(1 << 20) - 1, 12288, 24576, 49152, 98304]]).astype(object)
    xs = np.concatenate([np.arange(0, 300), rng.integers(0, 1 << 20, 4000), [(1 << 20) - 1]]).astype(object)
    for d in ds:
        m = (1 << 40) // int(d) + 1
        for x in (0, 1, int(d) - 1, int(d), int(d) + 1, 2 * int(d) - 1, (1 << 20) - 1):
            x = min(x, (1 << 20) - 1)
            assert (x * m) >> 40 == x // int(d), (x, d)
    for x, d in zip(xs, ds[:len(xs)]):
        m = (1 << 40) // int(d) + 1
        assert (int(x) * m) >> 40 == int(x) // int(d), (x, d)
