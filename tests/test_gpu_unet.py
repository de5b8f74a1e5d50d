"""GPU parity of the Attention-UNet engine against the torch-CPU fp32 oracle.

Tolerances (BASELINE.json north_star): probabilities max-abs 2e-2, masks >= 99.9 %
equal.  The oracle is torch fp32 of the published topology, NOT real onnxruntime
(absent offline) — see oracle/attunet_torch.py.
"""
import numpy as np
import pytest
import torch

from oracle.attunet_torch import oracle_unet_forward
from stroke_derenderer_b200.engine import UNetEngine
from stroke_derenderer_b200.synth import synth_line

pytestmark = pytest.mark.gpu

PROB_TOL = 2e-2
MASK_MIN_AGREE = 0.999


def _tiles(n, seed=0):
    """n tile inputs (B,3,128,384) f32 in [0,1]: synthetic handwriting cut like the reference does."""
    from oracle import segmentation_ref as O
    line = synth_line(320 * n + 200, seed)
    stack, *_ = O.cut_and_stack([line], (1, 3, 128, 384), 64)
    return (stack[:n] / 255.).astype(np.float32)


@pytest.fixture(scope="module")
def engine(cuda_device, parity_state):
    e = UNetEngine(parity_state, device=0, max_tiles=8, impl=0)
    yield e
    e.close()


def _compare(prob, ref, what, mask_min=MASK_MIN_AGREE):
    err = float(np.abs(prob - ref).max())
    agree = float(((prob > 0.5) == (ref > 0.5)).mean())
    print(f"[{what}] prob max-abs err {err:.5f}  mask agreement {agree * 100:.4f}%  fg(ref) {(ref > 0.5).mean() * 100:.2f}%")
    assert err <= PROB_TOL, (what, err)
    assert agree >= mask_min, (what, agree)


def test_empty_minibatch(engine):
    out = engine.run(None, {"input": np.zeros((0, 3, 128, 384), np.float32)})[0]
    assert out.shape == (0, 1, 128, 384)


@pytest.mark.parametrize("n", [1, 3, 8])
def test_tiles_vs_oracle(engine, oracle_net, n):
    x = _tiles(n, seed=40 + n)
    ref = oracle_unet_forward(oracle_net, x)
    prob = engine.run(None, {"input": x})[0]
    _compare(prob, ref, f"tiles{n}")


def test_intermediate_activations(engine, oracle_net):
    """Layer-by-layer check of the tcgen05 path (localises a bad descriptor / epilogue)."""
    x = _tiles(2, seed=7)
    taps = {}
    with torch.no_grad():
        oracle_net(torch.from_numpy(x), taps=taps)
    xt = UNetEngine.pack_input(torch.from_numpy(x).cuda())
    engine.forward(xt, want_prob32=True, want_mask=False)
    names = {"x1": "x1", "x2": "x2", "x3": "x3", "x4": "x4", "x5": "x5", "d5u": None, "a4": "a4", "d5": "d5",
             "a3": "a3", "d4": "d4", "a2": "a2", "d3": "d3", "a1": "a1"}
    worst = 0.0
    for tap, ref_name in names.items():
        if ref_name is None:
            continue
        got = engine.read_tap(tap, 2).float().cpu().numpy().transpose(0, 3, 1, 2)
        ref = taps[ref_name].numpy()
        rel = float(np.linalg.norm(got - ref) / (np.linalg.norm(ref) + 1e-12))
        print(f"[tap {tap}] rel-L2 err {rel:.5f}  max|ref| {np.abs(ref).max():.3f}  max-abs err {np.abs(got - ref).max():.4f}")
        worst = max(worst, rel)
    assert worst < 2e-2, worst


def test_simt_debug_impl_agrees(cuda_device, parity_state, oracle_net):
    e = UNetEngine(parity_state, device=0, max_tiles=2, impl=1)
    try:
        x = _tiles(2, seed=3)
        _compare(e.run(None, {"input": x})[0], oracle_unet_forward(oracle_net, x), "simt-debug")
    finally:
        e.close()


def test_binarize_images_vs_reference_golden(cuda_device, parity_state, golden_arrays):
    from stroke_derenderer_b200.evaluate_binarize import BinarizationSession
    bs = BinarizationSession()
    ort = bs.init_onnx_inference(parity_state)
    try:
        line = synth_line(1000, 9)
        out = bs.binarize_image(line, ort)
        assert out.shape == (128, 1000, 1) and out.dtype == np.uint8 and set(np.unique(out)) <= {0, 255}
        ref = np.unpackbits(golden_arrays["line1000_s9_binarized"])[:128 * 1000].reshape(128, 1000)
        agree = float(((out[:, :, 0] > 127) == (ref > 0)).mean())
        print(f"[binarize_image] mask agreement {agree * 100:.4f}%")
        assert agree >= MASK_MIN_AGREE
        # step-wise API gives the same result as the fused path
        stack, idx, widths, iw = bs.preprocess_images([line])
        step = bs.postprocess_stack(bs.model_predict(stack, ort), idx, widths, iw)[0]
        assert np.array_equal(step, out)
    finally:
        ort.close()


def test_fused_glue_equals_glue_of_tile_masks(cuda_device, parity_state):
    """sd_unet_forward_lines (head ORs into the zeroed line planes) == sd_unet_forward masks + sd_glue_u8, bit for
    bit, incl. 1-px and sub-tile lines, the 2-tile / 3-tile boundaries and the W mod n > 64 clipped tail (21000)."""
    from stroke_derenderer_b200 import segment as S
    widths = [1, 100, 383, 384, 385, 639, 640, 1000, 3072, 21000]
    lines = [synth_line(w, seed=70 + i) if w >= 40 else np.full((128, w, 3), 30, np.uint8) for i, w in enumerate(widths)]
    e = UNetEngine(parity_state, device=0, max_tiles=32, impl=0)
    try:
        seg = S.Segmenter(e)
        batch, planes = seg.binarize(lines)
        d_rgb = S.pack_lines_rgb(lines, batch).to(e.device)
        tiles = S.tile_extract_f16(batch, d_rgb)
        masks = torch.cat([e.forward(tiles[s:s + 32], want_mask=True)["mask"] for s in range(0, batch.n_tiles, 32)])
        ref = S.glue_u8(batch, masks.contiguous())
        assert bool((planes == ref).all())
        assert bool((planes != 0).any())
    finally:
        e.close()


def test_engine_switches_are_bit_identical(cuda_device, parity_state, monkeypatch):
    """The engine switches read at creation select different kernels for the same arithmetic: the psi-only level-1 gate whose
    consumer scales the skip rows (default) vs the gate that writes x * psi (SD_PSI_FUSED=0), and the level-2 layers on SM
    pairs (SD_CTA2_N128=1) vs the single-CTA form.  Probabilities and masks must agree bit for bit, on full and partial
    batches, and the on-demand `a1` tap of the psi-only engine must equal the tensor the other engine wrote."""
    x = UNetEngine.pack_input(torch.from_numpy(_tiles(9, seed=5)).cuda())

    def run(env):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        e = UNetEngine(parity_state, device=0, max_tiles=9, impl=0)
        try:
            full = e.forward(x, want_prob32=True, want_mask=True)
            out = [full["prob32"].clone(), full["mask"].clone()]
            part = e.forward(x[:4].contiguous(), want_prob32=True)
            out.append(part["prob32"].clone())
            out.append(e.read_tap("a1", 4).clone())
        finally:
            e.close()
            for k in env:
                monkeypatch.delenv(k, raising=False)
        return out

    base = run({})
    for env in ({"SD_PSI_FUSED": "0"}, {"SD_CTA2_N128": "1"}, {"SD_PSI_FUSED": "0", "SD_CTA2_N128": "1"}):
        other = run(env)
        for a, b in zip(base, other):
            assert torch.equal(a, b), env


def test_batch_invariance(cuda_device, parity_state):
    """Tiles are independent images: a tile's probabilities must not depend on what else is in the batch or on
    its position in it (catches cross-tile leaks in the persistent kernels: accumulator rings, pooled row pairs,
    phantom M tiles of odd groups).  Bit-exact, because every tile is reduced in the same order."""
    e = UNetEngine(parity_state, device=0, max_tiles=7, impl=0)
    try:
        x = _tiles(7, seed=11)
        xt = UNetEngine.pack_input(torch.from_numpy(x).cuda())
        full = e.forward(xt, want_prob32=True, want_mask=True)
        p_full, m_full = full["prob32"].cpu().numpy(), full["mask"].cpu().numpy()
        for idx in ([0], [6], [3, 1], [2, 4, 5], [6, 5, 4, 3, 2]):
            sub = e.forward(xt[idx].contiguous(), want_prob32=True, want_mask=True)
            assert np.array_equal(sub["prob32"].cpu().numpy(), p_full[idx]), idx
            assert np.array_equal(sub["mask"].cpu().numpy(), m_full[idx]), idx
    finally:
        e.close()


def test_large_batch_matches_small_batches(cuda_device, parity_state):
    """The benchmark's engine size (256 tiles per pass: offsets beyond 2^31 bytes in the level-1 tensors, the
    largest work-item counts) gives the same masks, bit for bit, as the same tiles through a 64-tile engine."""
    rng = np.random.default_rng(17)
    x = torch.from_numpy(rng.random((256, 128, 384, 8), dtype=np.float32)).cuda().half()
    x[..., 3:] = 0
    big = UNetEngine(parity_state, device=0, max_tiles=256, impl=0)
    try:
        m_big = big.forward(x, want_mask=True)["mask"].clone()
        m_tail = big.forward(x[:130].contiguous(), want_mask=True)["mask"].clone()    # a partial pass
    finally:
        big.close()
    small = UNetEngine(parity_state, device=0, max_tiles=64, impl=0)
    try:
        m_small = torch.cat([small.forward(x[s:s + 64].contiguous(), want_mask=True)["mask"].clone() for s in range(0, 256, 64)])
    finally:
        small.close()
    assert bool((m_big != 0).any()) and bool((m_big == 0).any())
    assert bool((m_big == m_small).all()) and bool((m_tail == m_small[:130]).all())


def test_repeatability_under_load(cuda_device, parity_state):
    """The same batch through the engine 12 times, back to back (persistent kernels, mbarrier pipelines, TMEM
    rings, cluster pairs): every run must reproduce the first bit for bit."""
    e = UNetEngine(parity_state, device=0, max_tiles=48, impl=0)
    try:
        rng = np.random.default_rng(3)
        x = torch.from_numpy(rng.random((48, 128, 384, 8), dtype=np.float32)).cuda().half()
        x[..., 3:] = 0
        first = e.forward(x, want_prob32=True, want_mask=True)
        p0, m0 = first["prob32"].clone(), first["mask"].clone()
        for _ in range(12):
            out = e.forward(x, want_prob32=True, want_mask=True)
            assert bool((out["prob32"] == p0).all()) and bool((out["mask"] == m0).all())
    finally:
        e.close()


def test_pipeline_job_equals_direct_segmentation(cuda_device, parity_state):
    """LineSegmentationJob (chunks, three streams, UNet batches that cross chunk boundaries, pinned D2H) returns
    exactly what one direct Segmenter pass over the same lines returns."""
    from stroke_derenderer_b200.pipeline import LineSegmentationJob
    from stroke_derenderer_b200.segment import Segmenter
    widths = [700, 1536, 333, 2048, 4000, 384, 1000, 3072, 640, 2222, 1234]
    lines = [synth_line(w, seed=50 + i) for i, w in enumerate(widths)]
    e = UNetEngine(parity_state, device=0, max_tiles=16, impl=0)
    try:
        ref = Segmenter(e).segment(lines)
        ref_parts = Segmenter(e).partition(ref["batch"], ref["planes"], canvases="device", crops=True)
        job = LineSegmentationJob(e, lines, lines_per_chunk=4)
        for results in (job.resident_step(), job.host_step()):
            torch.cuda.synchronize()
            li = 0
            for ch, res in zip(job.chunks, results):
                for k in range(ch.batch.n_lines):
                    assert torch.equal(ch.batch.plane(ch.planes, k), ref["batch"].plane(ref["planes"], li)), li
                    assert torch.equal(ch.batch.plane(res["labels"], k), ref["batch"].plane(ref["labels"], li)), li
                    a, b = int(res["line_group_start"][k]), int(res["line_group_start"][k + 1])
                    ra, rb = int(ref_parts["line_group_start"][li]), int(ref_parts["line_group_start"][li + 1])
                    assert b - a == rb - ra and np.array_equal(res["groups"][a:b, 1:5], ref_parts["groups"][ra:rb, 1:5]), li
                    assert torch.equal(res["crops"]["image"][a:b], ref_parts["crops"]["image"][ra:rb]), li
                    if "image_host" in res["crops"]:
                        assert np.array_equal(res["crops"]["image_host"][a:b], ref_parts["crops"]["image"][ra:rb].cpu().numpy()), li
                    li += 1
            assert li == len(lines)
    finally:
        e.close()


def test_api_calls_from_numpy_and_gather_arena(cuda_device, parity_state):
    """The reference-signature calls from plain numpy lists (`binarize_images`, `get_partitions_batch`), the fused
    `segment_lines` call and the shared-memory gather arena all return what one direct Segmenter pass returns."""
    import os
    from stroke_derenderer_b200 import gather as G
    from stroke_derenderer_b200.evaluate_binarize import BinarizationSession
    from stroke_derenderer_b200.evaluate_strokes import StrokeEstimationSession
    from stroke_derenderer_b200.pipeline import LineSegmentationJob, segment_lines
    from stroke_derenderer_b200.segment import Segmenter
    widths = [700, 1536, 333, 2048, 4000, 384, 1000, 3072, 640, 2222, 1234]
    lines = [synth_line(w, seed=150 + i) for i, w in enumerate(widths)]
    bs = BinarizationSession(max_tiles=16, lines_per_chunk=4)
    e = bs.init_onnx_inference(parity_state)
    try:
        ref = Segmenter(e).segment(lines)
        masks = bs.binarize_images(lines, e)
        for i, m in enumerate(masks):
            assert m.shape == (128, widths[i], 1) and m.dtype == np.uint8
            assert np.array_equal(m[:, :, 0], ref["batch"].plane(ref["planes"], i).cpu().numpy()), i
        se = StrokeEstimationSession()
        parts = se.get_partitions_batch([m[:, :, 0] > 127 for m in masks], lines_per_chunk=3)
        # the host lanes (threads + streams taking the chunks in turn) only change the schedule, never the result or its order
        parts1 = se.get_partitions_batch([m[:, :, 0] > 127 for m in masks], lines_per_chunk=2, lanes=1)
        assert [len(a) for a in parts] == [len(a) for a in parts1]
        for a, b in zip(parts, parts1):
            for pa, pb in zip(a, b):
                assert np.array_equal(pa["image"], pb["image"]) and pa["translate1"] == pb["translate1"] and pa["ratio"] == pb["ratio"]
        m2, p2 = segment_lines(e, lines, lines_per_chunk=4)
        assert all(np.array_equal(a, b) for a, b in zip(masks, m2))
        from oracle import segmentation_ref as O
        for i in range(len(lines)):
            want = O.get_partitions((masks[i][:, :, 0] > 127).astype(np.uint8))
            for got in (parts[i], p2[i]):
                assert len(got) == len(want), i
                for p, w in zip(got, want):
                    assert np.array_equal(p["image"], w["image"]) and np.array_equal(p["image_input"], w["image_input"])
                    assert (int(p["translate1"][0]), int(p["translate1"][1])) == (int(w["translate1"][0]), int(w["translate1"][1]))
                    assert p["ratio"] == w["ratio"] and tuple(p["translate2"]) == tuple(w["translate2"])
        # gather arena, one rank: D2H lands in the (page-locked) shared-memory region, the reader sees input order
        job = LineSegmentationJob(e, lines, lines_per_chunk=4, prepack=False)
        arena = G.ResultArena(f"sd_test_{os.getpid()}", [job.arena_bytes()], rank=0, create=True)
        try:
            arena.register()
            wr = G.RegionWriter(arena.region(0), len(job.chunks))
            for step in (1, 2):
                job.host_step(wr)
                got = G.GatheredResults(arena, [list(range(len(lines)))], widths, 4, step=step)
                for i in range(len(lines)):
                    assert np.array_equal(got.mask(i), masks[i][:, :, 0]), i
                    assert got.num(i) == int(ref["num"][i])
                    assert np.array_equal(got.crops(i), np.stack([p["image"] for p in p2[i]]) if len(p2[i]) else np.zeros((0, 224, 224), np.uint8))
                    cv = got.canvases(i)
                    rc = ref.line_canvases(i)
                    assert len(cv) == len(rc) and all(np.array_equal(a[0], b[0]) and a[1] == b[1] for a, b in zip(cv, rc))
            del got, cv
        finally:
            arena.close()
    finally:
        e.close()


def test_main_cli_dropin(cuda_device, parity_state, tmp_path):
    """B1 boundary: the root main.py (initialize_sessions / load_images / main) with a `binarizer.onnx` read by the
    dependency-free reader, PNG input, `_BINARIZED.png` + `_PARTITIONS.json` output."""
    import json
    import sys
    import cv2
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    sys.path.insert(0, str(root))
    import main as cli
    from onnx_writer import write_onnx as _write_onnx
    models, inp, out = tmp_path / "models", tmp_path / "in", tmp_path / "out"
    models.mkdir(); inp.mkdir()
    _write_onnx(models / "binarizer.onnx", parity_state, folded=True)
    (models / "configs_binarizer.json").write_text(json.dumps({"bin_thr": 0.5, "minibatch": 8}))
    lines = {"a": synth_line(900, 21), "b": synth_line(400, 22)}
    for stem, img in lines.items():
        cv2.imwrite(str(inp / f"{stem}.png"), cv2.cvtColor(img, cv2.COLOR_RGB2BGR))
    args = cli.parse_args(["-models", str(models), "-input", str(inp), "--output", str(out)])
    sessions = cli.initialize_sessions(args.models)
    try:
        imgs = cli.load_images(sorted(str(p) for p in Path(args.input).glob("*.png")))
        cli.main(imgs, *sessions, args.output, strokes=True)
        for stem, img in lines.items():
            png = cv2.imread(str(out / f"{stem}_BINARIZED.png"), cv2.IMREAD_GRAYSCALE)
            assert png is not None and png.shape == (128, img.shape[1]) and set(np.unique(png)) <= {0, 255}
            direct = sessions[0].binarize_image(img, sessions[1])[:, :, 0]
            assert np.array_equal(png, direct)
            parts = json.loads((out / f"{stem}_PARTITIONS.json").read_text())
            want = sessions[2].get_partitions(direct > 127)
            assert len(parts) == len(want)
            for p, w in zip(parts, want):
                assert p["translate1"] == [int(w["translate1"][0]), int(w["translate1"][1])] and p["ratio"] == float(w["ratio"])
    finally:
        sessions[1].close()


def test_main_cli_sharded_two_processes(cuda_device, parity_state, tmp_path):
    """The multi-GPU form of the CLI: `torchrun --nproc-per-node 2 main.py ...` (two rank processes; on a one-GPU box both
    use cuda:0).  Lines are dealt to the ranks, results come back through the shared-memory gather arena, rank 0 writes
    the files: they must equal what the single-process CLI writes."""
    import json
    import socket
    import subprocess
    import sys
    import cv2
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    sys.path.insert(0, str(root))
    import main as cli
    from onnx_writer import write_onnx as _write_onnx
    models, inp, out1, out2 = tmp_path / "models", tmp_path / "in", tmp_path / "out1", tmp_path / "out2"
    models.mkdir(); inp.mkdir()
    _write_onnx(models / "binarizer.onnx", parity_state, folded=True)
    (models / "configs_binarizer.json").write_text(json.dumps({"bin_thr": 0.5, "lines_per_chunk": 2, "max_tiles": 16}))
    widths = [900, 400, 1700, 2600, 640, 3100, 1234]
    for i, w in enumerate(widths):
        cv2.imwrite(str(inp / f"line{i:02d}.png"), cv2.cvtColor(synth_line(w, 600 + i), cv2.COLOR_RGB2BGR))
    sessions = cli.initialize_sessions(str(models))
    try:
        cli.main(cli.load_images(sorted(str(p) for p in inp.glob("*.png"))), *sessions, str(out1), strokes=True)
    finally:
        sessions[1].close()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(root / "main.py"), "-models", str(models), "-input", str(inp), "--output", str(out2)],
                       capture_output=True, text=True, timeout=900, cwd=str(root))
    assert r.returncode == 0, r.stderr[-3000:]
    names = sorted(p.name for p in out1.iterdir())
    assert names == sorted(p.name for p in out2.iterdir()) and len(names) == 2 * len(widths)
    for n in names:
        if n.endswith(".png"):
            assert np.array_equal(cv2.imread(str(out1 / n), cv2.IMREAD_UNCHANGED), cv2.imread(str(out2 / n), cv2.IMREAD_UNCHANGED)), n
        else:
            assert json.loads((out1 / n).read_text()) == json.loads((out2 / n).read_text()), n


def test_edge_case_lines_through_the_public_calls(cuda_device, parity_state):
    """Degenerate inputs through the pipelined public calls: no images, a line with no ink (no islands, no groups, no
    crops), a 1-px-wide line, a line of another height, and a chunk that consists of empty lines only — through
    binarize_images, segment_lines and the gather arena."""
    import os
    from oracle import segmentation_ref as O
    from stroke_derenderer_b200 import gather as G
    from stroke_derenderer_b200.evaluate_binarize import BinarizationSession
    from stroke_derenderer_b200.pipeline import LineSegmentationJob, segment_lines
    bs = BinarizationSession(max_tiles=16, lines_per_chunk=2)
    e = bs.init_onnx_inference(parity_state)
    try:
        assert bs.binarize_images([], e) == []
        white = np.full((128, 700, 3), 255, np.uint8)
        lines = [white, white.copy(), synth_line(900, 811), np.full((128, 1, 3), 255, np.uint8), synth_line(500, 812, height=96), white[:, :40].copy()]
        masks, parts = segment_lines(e, lines, lines_per_chunk=2)
        assert len(masks) == len(parts) == len(lines)
        ort = O.TorchOrtSession(parity_state)
        ref_bs = O.BinarizationSessionRef()
        for i, (m, p) in enumerate(zip(masks, parts)):
            ref = ref_bs.binarize_image(lines[i], ort)
            assert m.shape == ref.shape, (i, m.shape, ref.shape)
            assert float((m == ref).mean()) >= 0.998, i
            want = O.get_partitions((m[:, :, 0] > 127).astype(np.uint8))
            assert len(p) == len(want), i
            for a, b in zip(p, want):
                assert np.array_equal(a["image"], b["image"]) and a["ratio"] == b["ratio"]
        assert len(parts[2]) > 0
        # the same lines through the gather arena
        job = LineSegmentationJob(e, lines, lines_per_chunk=2, prepack=False)
        widths = [m.shape[1] for m in masks]
        arena = G.ResultArena(f"sd_test_edge_{os.getpid()}", [job.arena_bytes()], rank=0, create=True)
        try:
            arena.register()
            wr = G.RegionWriter(arena.region(0), len(job.chunks))
            job.host_step(wr)
            got = G.GatheredResults(arena, [list(range(len(lines)))], widths, 2, step=1)
            for i in range(len(lines)):
                assert np.array_equal(got.mask(i), masks[i][:, :, 0]) and got.crops(i).shape[0] == len(parts[i]) == len(got.groups(i))
                assert got.stats(i).shape[0] == got.num(i) - 1
            del got
        finally:
            arena.close()
    finally:
        e.close()
