"""GPU parity of the bandwidth-bound stages against the oracle (bit-exact)."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import segmentation_ref as O
from stroke_derenderer_b200 import segment as S
from stroke_derenderer_b200.synth import config_widths, ink_mask, synth_dense_mask, synth_line

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _batch(images, dev=0):
    device = torch.device("cuda", dev)
    batch = S.plan_batch([im.shape[1] for im in images], device)
    d_rgb = S.pack_lines_rgb(images, batch).to(device)
    return batch, d_rgb


WIDTHS = [1, 2, 100, 383, 384, 385, 639, 640, 1000, 1536, 3072, 3840, 6144, 16384, 20480, 21000]


def test_tile_extract_u8_matches_reference_golden(cuda_device, golden):
    for W in WIDTHS:
        img = np.random.default_rng(W).integers(0, 256, (128, W, 3), dtype=np.uint8)
        batch, d_rgb = _batch([img])
        stack = S.tile_extract_u8(batch, d_rgb).cpu().numpy()
        g = golden["geometry"][str(W)]
        assert stack.shape[0] == g["n"] and batch.stack_widths()[0] == g["widths"], W
        assert sha(stack) == g["stack_sha"], W


def test_tile_extract_batch_vs_oracle(cuda_device):
    rng = np.random.default_rng(3)
    imgs = [rng.integers(0, 256, (128, int(w), 3), dtype=np.uint8) for w in [300, 1000, 384, 5, 2222, 777]]
    batch, d_rgb = _batch(imgs)
    stack = S.tile_extract_u8(batch, d_rgb).cpu().numpy()
    ref, idx, widths, iw = O.cut_and_stack(imgs, (1, 3, 128, 384), 64)
    assert np.array_equal(stack, ref)
    assert batch.stack_indices() == idx and batch.stack_widths() == widths
    # fp16 NHWC8 form == np.float16(np.float32(x / 255.)) of the same cut, channels 3..7 zero
    t16 = S.tile_extract_f16(batch, d_rgb).cpu().numpy()
    want = (ref / 255.).astype(np.float32).astype(np.float16).transpose(0, 2, 3, 1)
    assert np.array_equal(t16[..., :3], want)
    assert not t16[..., 3:].any()


def test_u8_to_half_all_values(cuda_device):
    img = np.zeros((128, 256, 3), np.uint8)
    img[:, :, 0] = np.arange(256)[None, :]
    img[:, :, 1] = np.arange(256)[None, ::-1]
    img[:, :, 2] = 7
    batch, d_rgb = _batch([img])
    t16 = S.tile_extract_f16(batch, d_rgb).cpu().numpy()
    want = (img[:, :, :] / 255.).astype(np.float32).astype(np.float16)
    assert np.array_equal(t16[0, :, :256, :3], want)


def test_glue_u8_matches_reference_golden(cuda_device, golden):
    for W in WIDTHS:
        g = golden["geometry"][str(W)]
        out = (np.random.default_rng(W + 1).random((g["n"], 1, 128, 384)) < 0.3).astype(np.uint8) * 255
        batch = S.plan_batch([W], torch.device("cuda", 0))
        planes = S.glue_u8(batch, torch.from_numpy(out[:, 0]).cuda().contiguous())
        glued = batch.plane(planes, 0).cpu().numpy()
        assert glued.shape == (128, W)
        assert int(glued.sum(dtype=np.int64)) == g["glue_sum"], W
        assert sha(glued[:, :, None]) == g["glue_sha"], W
        # pitch padding must be zero (CCL relies on it)
        ln = batch.lines[0]
        full = planes[:128 * int(ln["pitch"])].view(128, int(ln["pitch"])).cpu().numpy()
        assert not full[:, W:].any()


def test_glue_random_values_and_threshold(cuda_device):
    rng = np.random.default_rng(11)
    widths = [300, 1000, 384, 4321, 640]
    batch = S.plan_batch(widths, torch.device("cuda", 0))
    n = batch.n_tiles
    vals = rng.integers(0, 256, (n, 1, 128, 384), dtype=np.uint8)        # arbitrary u8 -> true max semantics
    planes = S.glue_u8(batch, torch.from_numpy(vals[:, 0]).cuda().contiguous())
    ref = O.reconstruct_images(vals, widths, batch.stack_indices(), batch.stack_widths(), 64)
    for i in range(len(widths)):
        assert np.array_equal(batch.plane(planes, i).cpu().numpy(), ref[i][:, :, 0]), widths[i]
    prob = rng.random((n, 128, 384)).astype(np.float16)
    prob[0, 0, :8] = [0.5, 0.5005, 0.4995, 0.0, 1.0, 0.25, 0.75, 0.5]
    pl = S.glue_threshold_f16(batch, torch.from_numpy(prob).cuda(), 0.5, 255)
    binv = (255 * (prob.astype(np.float32) > 0.5)).astype(np.uint8)[:, None]
    ref = O.reconstruct_images(binv, widths, batch.stack_indices(), batch.stack_widths(), 64)
    for i in range(len(widths)):
        assert np.array_equal(batch.plane(pl, i).cpu().numpy(), ref[i][:, :, 0]), widths[i]


def test_glue_other_overlaps_and_paste_table(cuda_device):
    """helper/split.py accepts any 0 <= overlap < tile_w.  Overlaps up to the tile stride are glued bit-exactly; beyond
    that reconstruct_images pastes clipped inner tiles at shifted positions (s += width - overlap), which the gather
    form does not reproduce: sd_plan_lines refuses those geometries instead of returning a different mask."""
    from stroke_derenderer_b200 import _lib
    rng = np.random.default_rng(23)
    for overlap in (0, 64, 100, 150):
        widths = [384, 500, 1000, 2000, 777]
        batch = S.plan_batch(widths, torch.device("cuda", 0), overlap=overlap)
        vals = rng.integers(0, 256, (batch.n_tiles, 1, 128, 384), dtype=np.uint8)
        planes = S.glue_u8(batch, torch.from_numpy(vals[:, 0]).cuda().contiguous())
        ref = O.reconstruct_images(vals, widths, batch.stack_indices(), batch.stack_widths(), overlap)
        for i, w in enumerate(widths):
            assert ref[i].shape[1] == w
            assert np.array_equal(batch.plane(planes, i).cpu().numpy(), ref[i][:, :, 0]), (overlap, w)
        # the paste table of the fused head names the same positions
        tab = _lib.tile_dst_table(batch.lines, 0)
        k = 0
        for ln, ws in zip(batch.lines, batch.stack_widths()):
            for i, wd in enumerate(ws):
                start = 0 if int(ln["n_tiles"]) == 1 else i * int(ln["wu"])
                assert int(tab[k]["d_dst"]) == int(ln["px_off"]) + start and int(tab[k]["width"]) == min(wd, 384) and int(tab[k]["pitch"]) == int(ln["pitch"])
                k += 1
    for overlap in (200, 300, 380):
        with pytest.raises(_lib.SdError):
            S.plan_batch([384, 1000], torch.device("cuda", 0), overlap=overlap)


def test_cut_identity_glue_roundtrip_full_size(cuda_device):
    """Size-independent property at BASELINE config-3 scale: cut -> identity on channel 0 -> glue == input."""
    widths = config_widths(64)
    rng = np.random.default_rng(5)
    imgs = [rng.integers(0, 256, (128, int(w), 3), dtype=np.uint8) for w in widths]
    batch, d_rgb = _batch(imgs)
    stack = S.tile_extract_u8(batch, d_rgb)
    planes = S.glue_u8(batch, stack[:, 0].contiguous())
    for i, im in enumerate(imgs):
        assert np.array_equal(batch.plane(planes, i).cpu().numpy(), im[:, :, 0])


def _pack_masks(masks, dev=0):
    device = torch.device("cuda", dev)
    batch = S.plan_batch([m.shape[1] for m in masks], device)
    host = np.zeros(batch.px_total, np.uint8)
    for m, ln in zip(masks, batch.lines):
        off, pitch = int(ln["px_off"]), int(ln["pitch"])
        host[off:off + 128 * pitch].reshape(128, pitch)[:, :m.shape[1]] = m
    return batch, torch.from_numpy(host).to(device)


def test_ccl_labels_match_opencv_golden(cuda_device, golden, golden_arrays):
    names = list(golden["islands"].keys())
    masks = []
    for nme in names:
        shp = golden["islands"][nme]["shape"]
        masks.append(np.unpackbits(golden_arrays[f"{nme}_mask"])[:shp[0] * shp[1]].reshape(shp))
    batch, planes = _pack_masks(masks)
    labels, num = S.ccl_label(batch, planes)
    num = num.cpu().numpy()
    for i, nme in enumerate(names):
        assert int(num[i]) == golden["islands"][nme]["num"], nme
        got = batch.plane(labels, i).cpu().numpy()
        assert np.array_equal(got, golden_arrays[f"{nme}_labels"]), nme


def test_ccl_random_masks_vs_cv2(cuda_device):
    import cv2
    rng = np.random.default_rng(0)
    masks = []
    for t in range(48):
        W = int(rng.integers(1, 1500))
        p = float(rng.choice([0.02, 0.1, 0.3, 0.5, 0.7]))
        m = (rng.random((128, W)) < p).astype(np.uint8)
        if t % 3 == 0:
            m = cv2.dilate(m, np.ones((2, 2), np.uint8))
        masks.append(m)
    masks += [np.zeros((128, 33), np.uint8), np.ones((128, 17), np.uint8), np.eye(128, dtype=np.uint8),
              np.fliplr(np.eye(128, dtype=np.uint8)).copy()]
    batch, planes = _pack_masks(masks)
    labels, num = S.ccl_label(batch, planes)
    num_h = num.cpu().numpy()
    stats, stat_off, _ = S.island_stats(batch, labels, num_h)
    stats = stats.cpu().numpy()
    for i, m in enumerate(masks):
        n, ref, st, _ = cv2.connectedComponentsWithStats(m)
        assert int(num_h[i]) == n, i
        assert np.array_equal(batch.plane(labels, i).cpu().numpy(), ref), i
        assert np.array_equal(stats[stat_off[i]:stat_off[i + 1]], st[1:]), i


def test_ccl_more_runs_than_shared_parents(cuda_device):
    """Strips with more than 2048 runs (every 2x2 block its own run: pixels on even columns only) keep their union-find
    parents in the global scratch array instead of shared memory; labels, counts and stats stay cv2's."""
    import cv2
    rng = np.random.default_rng(9)
    yy, xx = np.mgrid[0:128, 0:700]
    stripes = (xx % 2 == 0).astype(np.uint8)                                   # 64 runs per block row, 4096 per strip
    masks = [stripes, stripes * (yy % 4 < 2), stripes * (rng.random((128, 700)) < 0.7),
             np.maximum(stripes * (yy < 70), (rng.random((128, 700)) < 0.3) * (yy >= 60)).astype(np.uint8),
             ((xx + yy) % 2 == 0).astype(np.uint8), (xx % 4 == 0).astype(np.uint8)]
    masks = [np.ascontiguousarray(m, dtype=np.uint8) for m in masks]
    batch, planes = _pack_masks(masks)
    labels, meta, stats = S.ccl_label_stats(batch, planes, 400_000)
    n = batch.n_lines
    stat_off = meta[:8 * (n + 1)].view(torch.int64).cpu().numpy()
    num = meta[8 * (n + 1):].view(torch.int32).cpu().numpy()
    stats = stats.cpu().numpy()
    for i, m in enumerate(masks):
        cn, ref, st, _ = cv2.connectedComponentsWithStats(m)
        assert int(num[i]) == cn, i
        assert np.array_equal(batch.plane(labels, i).cpu().numpy(), ref), i
        assert np.array_equal(stats[stat_off[i]:stat_off[i + 1]], st[1:]), i


def test_ccl_build_switches_give_the_same_labels(cuda_device, monkeypatch):
    """The A/B switches of the CCL chain (label kernel compiled for 10 / 12 CTAs per SM, the latter with 1536 shared
    parents per strip; seam merges folded into the line kernel) are read per call: every variant returns cv2's labels,
    counts and stats, including on a strip whose runs no longer fit in the smaller parent table."""
    import cv2
    rng = np.random.default_rng(21)
    yy, xx = np.mgrid[0:128, 0:700]
    masks = [ink_mask(synth_line(3072, seed=77)), (rng.random((128, 1300)) < 0.3).astype(np.uint8),
             np.ascontiguousarray(((xx % 2 == 0) & (yy % 8 < 3)), dtype=np.uint8),        # 1 536 < runs per strip <= 2 048
             synth_dense_mask(4096, 0.01, 3)]
    refs = [cv2.connectedComponentsWithStats(m) for m in masks]
    batch, planes = _pack_masks(masks)
    for occ, merge in (("10", "0"), ("12", "0"), ("12", "1"), ("8", "1")):
        monkeypatch.setenv("SD_CCL_OCC", occ)
        monkeypatch.setenv("SD_CCL_MERGE", merge)
        labels, meta, stats = S.ccl_label_stats(batch, planes, 200_000)
        n = batch.n_lines
        meta_h = meta.cpu().numpy()
        stat_off, num = meta_h[:8 * (n + 1)].view(np.int64), meta_h[8 * (n + 1):].view(np.int32)
        st = stats.cpu().numpy()
        for i, (rn, rl, rs, _) in enumerate(refs):
            assert int(num[i]) == rn, (occ, merge, i)
            assert np.array_equal(batch.plane(labels, i).cpu().numpy(), rl), (occ, merge, i)
            assert np.array_equal(st[stat_off[i]:stat_off[i + 1]], rs[1:]), (occ, merge, i)


def test_ccl_dense_config5(cuda_device):
    import cv2
    masks = [synth_dense_mask(16384, 0.003, 0), synth_dense_mask(16384, 0.01, 1)]
    batch, planes = _pack_masks(masks)
    labels, num = S.ccl_label(batch, planes)
    for i, m in enumerate(masks):
        n, ref = cv2.connectedComponents(m)
        assert int(num[i].item()) == n
        assert np.array_equal(batch.plane(labels, i).cpu().numpy(), ref)


def test_ccl_label_stats_fused_vs_cv2(cuda_device):
    """sd_ccl_label_stats: labels AND the cv2 stats rows (x, y, w, h, area) from one pass, bit-exact against
    cv2.connectedComponentsWithStats on text-like, random, dense and degenerate masks; row offsets = exclusive scan
    of the island counts; a too-small row reserve is reported through stat_off[n_lines] (rows beyond it dropped)."""
    import cv2
    rng = np.random.default_rng(77)
    masks = [ink_mask(synth_line(int(w), seed=300 + i)) for i, w in enumerate([700, 3072, 129, 6144, 128, 2222])]
    for k in range(10):
        W = int(rng.integers(1, 1500))
        m = (rng.random((128, W)) < rng.choice([0.02, 0.2, 0.5, 0.8])).astype(np.uint8)
        masks.append(cv2.dilate(m, np.ones((2, 2), np.uint8)) if k % 2 else m)
    masks += [synth_dense_mask(4096, 0.01, 5), np.zeros((128, 300), np.uint8), np.ones((128, 257), np.uint8),
              np.eye(128, dtype=np.uint8), np.fliplr(np.eye(128, dtype=np.uint8))]
    batch, planes = _pack_masks(masks)
    for cap in (S.stats_capacity(batch) * 4, 100):
        labels, meta, stats = S.ccl_label_stats(batch, planes, cap)
        torch.cuda.synchronize()
        n = batch.n_lines
        meta_h = meta.cpu().numpy()
        stat_off, num = meta_h[:8 * (n + 1)].view(np.int64), meta_h[8 * (n + 1):].view(np.int32)
        st = stats.cpu().numpy()
        exp_off = 0
        for i, m in enumerate(masks):
            rn, rl, rs, _ = cv2.connectedComponentsWithStats(m)
            assert int(num[i]) == rn and int(stat_off[i]) == exp_off, i
            assert np.array_equal(batch.plane(labels, i).cpu().numpy(), rl), i
            a, b = exp_off, min(exp_off + rn - 1, cap)
            if b > a:
                assert np.array_equal(st[a:b], rs[1:1 + b - a]), (i, cap)
            exp_off += rn - 1
        assert int(stat_off[n]) == exp_off
        assert exp_off > 100                          # the second round really overflows its reserve


def test_partition_matches_reference_golden(cuda_device, golden, golden_arrays):
    from stroke_derenderer_b200.evaluate_strokes import StrokeEstimationSession
    names = [k for k in golden["islands"] if golden["islands"][k]["num"] > 1]
    masks = []
    for nme in names:
        shp = golden["islands"][nme]["shape"]
        masks.append(np.unpackbits(golden_arrays[f"{nme}_mask"])[:shp[0] * shp[1]].reshape(shp))
    se = StrokeEstimationSession()
    parts = se.get_partitions_batch(masks)
    batch, planes = _pack_masks(masks)
    res = S.Segmenter(None, device=torch.device("cuda", 0)).partition(batch, planes)
    for i, nme in enumerate(names):
        g = golden["islands"][nme]
        canv = res.line_canvases(i)
        assert len(canv) == len(g["groups"]), nme
        for (c, (top, left)), gg in zip(canv, g["groups"]):
            assert [int(top), int(left)] == gg["pos"] and list(c.shape) == gg["shape"] and sha(c) == gg["sha"], nme
        assert len(parts[i]) == len(g["partitions"])
        for p, gp in zip(parts[i], g["partitions"]):
            assert [int(p["translate1"][0]), int(p["translate1"][1])] == gp["translate1"]
            assert p["ratio"] == gp["ratio"] and list(p["translate2"]) == gp["translate2"]
            assert sha(p["image"]) == gp["image_sha"] and sha(p["image_input"]) == gp["input_sha"]


def test_get_binarized_islands_dropin(cuda_device, golden, golden_arrays):
    from stroke_derenderer_b200.helper.partition import get_binarized_islands, group_islands
    for nme in ["line300", "line1000", "long_island", "full40", "empty64"]:
        g = golden["islands"][nme]
        shp = g["shape"]
        m = np.unpackbits(golden_arrays[f"{nme}_mask"])[:shp[0] * shp[1]].reshape(shp)
        islands, lab, num = get_binarized_islands(m, margin=2)
        assert num == g["num"] and np.array_equal(lab, golden_arrays[f"{nme}_labels"])
        assert len(islands) == len(g["islands"])
        for (c, pos), gi in zip(islands, g["islands"]):
            assert list(pos) == gi["pos"] and list(c.shape) == gi["shape"] and sha(c) == gi["sha"]
        groups = group_islands(islands, (128, 128)) if islands else []
        assert len(groups) == len(g["groups"])
        for (c, pos), gg in zip(groups, g["groups"]):
            assert [int(pos[0]), int(pos[1])] == gg["pos"] and sha(c) == gg["sha"]


def test_split_module_dropin(cuda_device):
    from stroke_derenderer_b200.helper.split import cut_and_stack, reconstruct_images
    rng = np.random.default_rng(21)
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in [(128, 300), (128, 1000), (200, 1000)]]
    a = cut_and_stack(imgs, (1, 3, 128, 384), 64)
    b = O.cut_and_stack(imgs, (1, 3, 128, 384), 64)
    assert np.array_equal(a[0], b[0]) and a[1] == b[1] and a[2] == b[2] and a[3] == b[3]
    out = rng.integers(0, 2, (a[0].shape[0], 1, 128, 384)).astype(np.uint8) * 255
    ra = reconstruct_images(out, a[3], a[1], a[2], 64)
    rb = O.reconstruct_images(out, b[3], b[1], b[2], 64)
    assert all(np.array_equal(x, y) for x, y in zip(ra, rb))


def test_resize_lines_bit_exact_vs_cv2(cuda_device):
    """sd_resize_lines == resize_to_height (common.py:85-93: cv2.resize, INTER_LINEAR) on lines of other heights,
    mixed in one batch with 128-px lines: up- and down-scaling, the exact-2x area path (even and odd widths),
    1- and 2-row sources, widths that are not multiples of the 128-column CTA, and reductions beyond the
    kernel's staging window (direct-gather fallback)."""
    rng = np.random.default_rng(11)
    shapes = [(200, 1000), (128, 700), (64, 300), (256, 1024), (256, 1025), (130, 777), (127, 500), (1, 40), (2, 9),
              (128, 384), (129, 1290), (37, 400), (255, 1001), (512, 3000), (300, 2000), (90, 12000), (800, 5000), (1500, 4000), (679, 2037), (700, 700)]
    shapes += [(int(rng.integers(3, 400)), int(rng.integers(8, 1500))) for _ in range(20)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes if int(w * (128 / h)) >= 1]
    batch, d_rgb = S.upload_lines(imgs, torch.device("cuda", 0))
    got = d_rgb.cpu().numpy()
    for im, ln in zip(imgs, batch.lines):
        ref = O.resize_to_height(im, 128) if im.shape[0] != 128 else im
        assert int(ln["width"]) == ref.shape[1]
        off = int(ln["img_off"])
        assert np.array_equal(got[off:off + ref.size].reshape(ref.shape), ref), im.shape


def test_binarize_images_other_heights(cuda_device, parity_state):
    """BinarizationSession.binarize_images on lines taller / shorter than 128 (evaluate_binarize.py:76) == the same
    call on the lines resized by the reference's cv2 call first; and the chunked pipeline agrees."""
    from stroke_derenderer_b200.evaluate_binarize import BinarizationSession
    from stroke_derenderer_b200.pipeline import LineSegmentationJob
    bs = BinarizationSession(max_tiles=16)
    ort = bs.init_onnx_inference(parity_state)
    lines = [synth_line(900, 3), synth_line(1500, 4), synth_line(400, 5)]
    import cv2
    raw = [cv2.resize(lines[0], (1400, 200)), lines[1], cv2.resize(lines[2], (300, 96))]
    a = bs.binarize_images(raw, ort)
    b = bs.binarize_images([O.resize_to_height(im, 128) if im.shape[0] != 128 else im for im in raw], ort)
    assert [x.shape for x in a] == [x.shape for x in b]
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    job = LineSegmentationJob(ort, raw, lines_per_chunk=2, crops=False)
    for step in (job.resident_step, job.host_step):
        res = step()
        torch.cuda.synchronize()
        k = 0
        for ch in job.chunks:
            for i in range(ch.batch.n_lines):
                assert np.array_equal(ch.batch.plane(ch.planes, i).cpu().numpy(), a[k][:, :, 0]), (step.__name__, k)
                k += 1


def test_group_crops_bit_exact_vs_reference_calls(cuda_device):
    """sd_group_crops == normalize -> cv2.resize -> pad -> normalize -> mean/std of evaluate_strokes.py:202-222,
    on hand-made canvases that hit every branch: up- and down-scaling, exact 2x decimation (area path), constant
    canvas (normalises to zero), 1-pixel canvases, long islands."""
    import cv2
    rng = np.random.default_rng(7)
    shapes = [(100, 444), (1, 1), (128, 3), (2, 300), (64, 64), (111, 111), (128, 2000), (37, 222), (128, 128), (5, 7)]
    shapes += [(int(rng.integers(1, 129)), int(rng.integers(1, 600))) for _ in range(40)]
    canv = []
    for k, (h, w) in enumerate(shapes):
        c = (rng.random((h, w)) < (0.2 if k % 2 else 0.6)).astype(np.uint8)
        if k == 4:
            c[:] = 1                                   # constant canvas -> cv2.normalize gives zeros
        c[rng.integers(0, h), rng.integers(0, w)] = 1   # a group always holds at least one ink pixel
        canv.append(c)
    dev = torch.device("cuda", 0)
    table = np.zeros((len(canv), 6), np.int64)
    off = 0
    for g, c in enumerate(canv):
        table[g] = (0, 10, 3, 10 + c.shape[1], 3 + c.shape[0], off)
        off += c.size
    flat = torch.from_numpy(np.concatenate([c.reshape(-1) for c in canv])).to(dev)
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    out = S.group_crops(dev, flat, torch.from_numpy(table).to(dev), table, lut=S.input_lut(mean, std))
    img = out["image"].cpu().numpy(); inp = out["image_input"].cpu().numpy()
    for g, c in enumerate(canv):
        ref_img, ratio, t2 = O.resize_and_pad_image(O.normalize_image(c), (224, 224), margin=1, pad_value=0)
        assert np.array_equal(img[g], ref_img), (g, c.shape)
        assert np.array_equal(inp[g], O.model_input_from_crop(ref_img, mean, std)), (g, c.shape)
        assert out["ratio"][g] == ratio and tuple(out["translate2"][g]) == t2


def _properties_of_partition(batch, planes, res):
    """Size-independent invariants of labelling + grouping, checked on the device for every line of a batch."""
    labels = res["labels"]
    mask = planes != 0
    # 1. labels cover exactly the mask
    assert bool(((labels > 0) == mask).all())
    # 2. idempotence: labelling the labelled mask again gives the same labels (numbering is a function of the mask)
    labels2, num2 = S.ccl_label(batch, (labels > 0).to(torch.uint8))
    assert bool((labels2 == labels).all()) and np.array_equal(num2.cpu().numpy(), res["num"])
    # 3. stats: areas sum to the mask, every label 1..n-1 is used, boxes lie inside the image
    st, off = res["stats"], res["stat_off"]
    assert int(st[:, 4].sum()) == int(mask.sum().item())
    assert (st[:, 4] > 0).all() and (st[:, 0] >= 0).all() and (st[:, 1] >= 0).all() and (st[:, 1] + st[:, 3] <= 128).all()
    for l in range(batch.n_lines):
        assert (st[off[l]:off[l + 1], 0] + st[off[l]:off[l + 1], 2] <= batch.widths[l]).all()
    assert np.array_equal(np.diff(off), res["num"].astype(np.int64) - 1)
    # 4. every island belongs to exactly one group, and the group canvases hold exactly the ink of their members
    assert (res["group_of"][:int(off[-1])] >= 0).all()
    assert res["canvas"].max().item() <= 1
    area_by_group = np.bincount(res["group_of"][:int(off[-1])], weights=st[:, 4], minlength=len(res["groups"])).astype(np.int64)
    assert int(res["canvas"].sum(dtype=torch.int64).item()) == int(area_by_group.sum())
    g = res["groups"]
    canvas_h = None
    for k in np.linspace(0, len(g) - 1, 64).astype(int):            # a spread of groups, canvas by canvas
        _, left, top, right, bottom, o = (int(v) for v in g[k])
        n_px = (right - left) * (bottom - top)
        assert int(res["canvas"][o:o + n_px].sum(dtype=torch.int64).item()) == int(area_by_group[k]), k


def test_full_size_config3_properties(cuda_device):
    """BASELINE config 3 at full size (512 lines, 6 438 tiles worth of columns): the oracle takes minutes there, so
    the device result is checked through invariants; two lines are also compared with cv2 bit for bit."""
    import cv2
    widths = config_widths(512)
    masks = [ink_mask(synth_line(int(w), seed=i)) for i, w in enumerate(widths)]
    batch, planes = _pack_masks(masks)
    res = S.Segmenter(None, device=torch.device("cuda", 0)).partition(batch, planes, canvases="device")
    _properties_of_partition(batch, planes, res)
    for i in (0, 511):
        n, ref = cv2.connectedComponents(masks[i])
        assert int(res["num"][i]) == n and np.array_equal(batch.plane(res["labels"], i).cpu().numpy(), ref)


def test_full_size_config5_properties(cuda_device):
    """BASELINE config 5: 64 dense 128x16384 masks in one launch (~800 k islands)."""
    masks = [synth_dense_mask(16384, 0.003 if i % 2 == 0 else 0.01, seed=i) for i in range(64)]
    batch, planes = _pack_masks(masks)
    res = S.Segmenter(None, device=torch.device("cuda", 0)).partition(batch, planes, canvases="device")
    _properties_of_partition(batch, planes, res)


def test_ccl_union_race_regression(cuda_device):
    """A looped stroke whose two branches merge across a thread-word boundary: with path compression running
    concurrently with the unions this strip lost a link in ~25 % of launches (one component became two).  2 000
    copies in one launch make a rare race visible."""
    import cv2
    line = ink_mask(synth_line(int(config_widths(512)[127]), seed=127))
    strip = np.ascontiguousarray(line[:, 2560:2688])
    n_ref, ref = cv2.connectedComponents(strip)
    assert n_ref == 6                                  # the synthetic strip this test was written for
    batch, planes = _pack_masks([strip] * 2000)
    for _ in range(3):
        labels, num = S.ccl_label(batch, planes)
        assert bool((num == n_ref).all()), int((num != n_ref).sum().item())
        got = labels.view(2000, 128, 128)
        assert bool((got == torch.from_numpy(ref).to(got.device).to(torch.int32)[None]).all())


def test_stroke_front_end_encode_postprocess_and_batching(cuda_device):
    """_encode_postprocess (evaluate_strokes.py:72-91) on the GPU, bit-exact against the
    reference's numpy expression, and the batched front end: the device-resident crops of MANY lines go through an
    encoder handle in cross-line batches.  The encoder graph itself is not in the reference (Drive file, topology not
    even named), so the handle here is a small seeded torch module: the plumbing is what is checked."""
    from stroke_derenderer_b200.evaluate_strokes import StrokeEstimationSession
    se = StrokeEstimationSession()
    rng = np.random.default_rng(5)
    for B, C in ((1, 8), (5, 100), (33, 512)):
        enc = rng.standard_normal((B, C, 7, 7)).astype(np.float32)
        ref = np.zeros((B, C, 14, 14), np.float32)                 # the reference's expression, verbatim semantics
        ref[:, :, ::2, ::2] = enc; ref[:, :, 1::2, 1::2] = enc; ref[:, :, ::2, 1::2] = enc; ref[:, :, 1::2, ::2] = enc
        ref = np.reshape(np.transpose(ref, (0, 2, 3, 1)), (B, -1, C)).astype(np.float32)
        assert np.array_equal(se._encode_postprocess(enc), ref)
        assert torch.equal(se._encode_postprocess(torch.from_numpy(enc).cuda()).cpu(), torch.from_numpy(ref))

    class Encoder:                       # stand-in handle with both call forms; (B,3,224,224) -> (B,16,7,7)
        def __init__(self):
            torch.manual_seed(3)
            self.net = torch.nn.Sequential(torch.nn.Conv2d(3, 16, 7, 32, 3), torch.nn.ReLU()).cuda().eval()
            self.batches = []

        @torch.no_grad()
        def run_device(self, names, feeds):
            self.batches.append(int(feeds["input"].shape[0]))
            return [self.net(feeds["input"])]

        def run(self, names, feeds):
            return [self.run_device(names, {"input": torch.from_numpy(feeds["input"]).cuda()})[0].cpu().numpy()]

    masks = [ink_mask(synth_line(w, seed=500 + i)) for i, w in enumerate([900, 2000, 1300, 64])]
    enc_h = Encoder()
    orts = se.load_orts({"encoder": enc_h})
    parts, encs = se.encode_partitions_batch(masks, orts, max_batch=16)
    n = sum(len(p) for p in parts)
    assert n > 16 and sum(enc_h.batches) == n and max(enc_h.batches) == 16          # batches cross line boundaries
    for p, e in zip(parts, encs):
        assert e.shape[0] == len(p)
        if len(p):                       # per line == the reference order of operations on that line's own crops
            x = np.stack([q["image_input"] for q in p]).astype(np.float32)
            want = se._encode_postprocess(enc_h.run(["output"], {"input": x})[0])
            assert np.allclose(e.cpu().numpy(), want, rtol=1e-4, atol=1e-5) and e.shape[1:] == (196, 16)
    with pytest.raises(NotImplementedError):
        se.load_orts({"encoder": "models/encoder.onnx"})


def test_stroke_encoder_handle_vs_torch_fp32(cuda_device):
    """The library-path encoder handle (stride-32 ResNet trunk, recalled topology, seeded weights; cuDNN fp16
    channels-last) against the same module in torch-CPU fp32, and through the batched front end.  UNPINNED: the
    reference's encoder.onnx is a Drive file whose topology the tree does not name."""
    from stroke_derenderer_b200.evaluate_strokes import StrokeEstimationSession
    from stroke_derenderer_b200.stroke_encoder import EncoderHandle, seeded_trunk
    ref_net = seeded_trunk(18, seed=7)
    handle = EncoderHandle(seeded_trunk(18, seed=7), device=0, max_batch=8)
    se = StrokeEstimationSession()
    masks = [ink_mask(synth_line(w, seed=700 + i)) for i, w in enumerate([800, 1500])]
    parts, encs = se.encode_partitions_batch(masks, se.load_orts({"encoder": handle}), max_batch=8)
    n = sum(len(p) for p in parts)
    assert n >= 8
    x = np.stack([q["image_input"] for p in parts for q in p]).astype(np.float32)
    with torch.no_grad():
        want = ref_net(torch.from_numpy(x)).numpy()
    got_raw = handle.run(["output"], {"input": x})[0]
    assert got_raw.shape == want.shape == (n, 512, 7, 7)
    rel = float(np.linalg.norm(got_raw - want) / np.linalg.norm(want))
    print(f"[encoder handle] rel-L2 err vs torch fp32 {rel:.5f}")
    assert rel < 1e-2
    got = torch.cat([e for e in encs if e.shape[0]], 0).cpu().numpy()
    want_pp = se._encode_postprocess(want)
    assert got.shape == want_pp.shape == (n, 196, 512)
    assert float(np.linalg.norm(got - want_pp) / np.linalg.norm(want_pp)) < 1e-2
