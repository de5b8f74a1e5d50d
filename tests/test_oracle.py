"""Pins the oracle (oracle/segmentation_ref.py) two ways: against golden vectors
produced by running the reference (tests/golden/make_golden.py) and, where
/root/reference exists, live against the unmodified reference functions."""
import hashlib

import numpy as np
import pytest

from oracle import segmentation_ref as O
from stroke_derenderer_b200.synth import ink_mask, synth_dense_mask, synth_line


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _mask(golden, golden_arrays, name):
    shp = golden["islands"][name]["shape"]
    return np.unpackbits(golden_arrays[f"{name}_mask"])[:shp[0] * shp[1]].reshape(shp)


def test_tile_geometry_table():
    """SURVEY.md A.1 known answers."""
    table = {1: (1, 1, 1), 383: (1, 383, 383), 384: (2, 256, 192), 385: (2, 256, 193), 639: (2, 383, 320),
             640: (3, 277, 214), 1000: (4, 314, 250), 1536: (5, 371, 308), 3072: (10, 371, 309),
             3840: (13, 359, 300), 6144: (20, 371, 311), 16384: (52, 379, 319), 20480: (65, 379, 320),
             32768: (103, 382, 332)}
    for W, (n, first, last) in table.items():
        starts, widths = O.tile_geometry(W)
        assert (len(widths), widths[0], widths[-1]) == (n, first, last), W
        assert all(w <= 383 for w in widths) or W < 384


def test_split_glue_golden(golden):
    for W, g in golden["geometry"].items():
        W = int(W)
        img = np.random.default_rng(W).integers(0, 256, (128, W, 3), dtype=np.uint8)
        stack, idx, widths, iw = O.cut_and_stack([img], (1, 3, 128, 384), 64)
        assert stack.shape[0] == g["n"] and [int(w) for w in widths[0]] == g["widths"] and sha(stack) == g["stack_sha"]
        out = (np.random.default_rng(W + 1).random((stack.shape[0], 1, 128, 384)) < 0.3).astype(np.uint8) * 255
        glued = O.reconstruct_images(out, iw, idx, widths, 64)[0]
        assert sha(glued) == g["glue_sha"]


def test_multi_image_bookkeeping_golden(golden):
    imgs = [np.random.default_rng(10 + i).integers(0, 256, (h, w, 3), dtype=np.uint8)
            for i, (h, w) in enumerate([(128, 300), (128, 1000), (128, 384), (200, 1000)])]
    stack, idx, widths, iw = O.cut_and_stack(imgs, (1, 3, 128, 384), 64)
    g = golden["multi"]
    assert list(stack.shape) == g["shape"] and idx == g["indices"] and iw == g["img_widths"]
    assert [[int(x) for x in w] for w in widths] == g["widths"] and sha(stack) == g["stack_sha"]
    assert idx[:3] == [[0], [1, 2, 3, 4], [5, 6]] and iw[3] == 640


def test_cut_identity_glue_roundtrip():
    for W in [1, 383, 384, 385, 640, 1000, 3072, 16384]:
        img = np.random.default_rng(W).integers(0, 256, (128, W, 3), dtype=np.uint8)
        stack, idx, widths, iw = O.cut_and_stack([img], (1, 3, 128, 384), 64)
        back = O.reconstruct_images(stack[:, :1], iw, idx, widths, 64)[0]
        assert np.array_equal(back[:, :, 0], img[:, :, 0]), W


def test_group_intervals_golden(golden):
    for case in golden["group_intervals"]:
        iv = [tuple(x) for x in case["intervals"]]
        assert O.group_intervals(iv, 128) == case["groups"]


def test_islands_groups_partitions_golden(golden, golden_arrays):
    for name, g in golden["islands"].items():
        m = _mask(golden, golden_arrays, name)
        islands, labels, num = O.get_binarized_islands(m, 2)
        assert num == g["num"] and np.array_equal(labels, golden_arrays[f"{name}_labels"])
        assert [[int(p[0]), int(p[1])] for _, p in islands] == [i["pos"] for i in g["islands"]]
        assert [sha(c) for c, _ in islands] == [i["sha"] for i in g["islands"]]
        if not islands:
            continue
        groups = O.group_islands(islands, (128, 128))
        assert [[int(p[0]), int(p[1])] for _, p in groups] == [x["pos"] for x in g["groups"]]
        assert [sha(c) for c, _ in groups] == [x["sha"] for x in g["groups"]]
        parts = O.get_partitions(m)
        assert [sha(p["image_input"]) for p in parts] == [x["input_sha"] for x in g["partitions"]]
        assert [p["ratio"] for p in parts] == [x["ratio"] for x in g["partitions"]]


def test_ccl_numbering_rule(golden, golden_arrays):
    """OpenCV numbers components by the raster key of their first 2x2 block."""
    from scipy import ndimage
    for name in ["line1000", "dense2048", "long_island"]:
        m = _mask(golden, golden_arrays, name)
        ref = golden_arrays[f"{name}_labels"]
        comp, n = ndimage.label(m, structure=np.ones((3, 3)))
        H, W = m.shape
        rr, cc = np.nonzero(m)
        key = (rr // 2) * ((W + 1) // 2) + cc // 2
        kmin = np.full(n + 1, np.iinfo(np.int64).max)
        np.minimum.at(kmin, comp[rr, cc], key)
        rank = np.zeros(n + 1, np.int64)
        rank[1:][np.argsort(kmin[1:])] = np.arange(1, n + 1)
        assert np.array_equal(rank[comp], ref), name


def test_unet_oracle_golden(oracle_net, golden, golden_arrays):
    from oracle.attunet_torch import oracle_unet_forward
    x = np.random.default_rng(0).random((1, 3, 128, 384), dtype=np.float32)
    p = oracle_unet_forward(oracle_net, x)
    assert np.abs(p - golden_arrays["config1_prob"]).max() < 1e-4
    assert abs(float(p.mean()) - golden["unet"]["config1_prob_mean"]) < 1e-5


def test_oracle_session_semantics(oracle_net):
    """evaluate_binarize.py:93-106: empty last minibatch, strict >, {0,255} output."""
    class Handle:
        calls = []
        def run(self, names, feeds):
            x = feeds["input"]; Handle.calls.append(x.shape[0])
            assert x.dtype == np.float32 and (x.shape[0] == 0 or x.max() <= 1.0)
            return [np.full((x.shape[0], 1, 128, 384), 0.5, np.float32) if x.shape[0] else np.zeros((0, 1, 128, 384), np.float32)]
    bs = O.BinarizationSessionRef()
    out = bs.model_predict(np.zeros((8, 3, 128, 384), np.uint8), Handle())
    assert Handle.calls == [8, 0] and out.shape == (8, 1, 128, 384) and out.max() == 0   # 0.5 is NOT > 0.5


# ---- live against the reference (build container only) -------------------------------
def test_live_split_vs_reference(reference_modules):
    RS = reference_modules["split"]
    rng = np.random.default_rng(0)
    for W in [1, 383, 384, 385, 639, 640, 1000, 1536, 3072, 6144, 21000]:
        img = rng.integers(0, 256, (128, W, 3), dtype=np.uint8)
        a = RS.cut_and_stack([img], (1, 3, 128, 384), 64); b = O.cut_and_stack([img], (1, 3, 128, 384), 64)
        assert np.array_equal(a[0], b[0]) and a[1:] == b[1:]
        out = rng.integers(0, 256, (a[0].shape[0], 1, 128, 384), dtype=np.uint8)
        ra = RS.reconstruct_images(out, a[3], a[1], a[2], 64); rb = O.reconstruct_images(out, b[3], b[1], b[2], 64)
        assert all(np.array_equal(x, y) for x, y in zip(ra, rb))


def test_live_partition_vs_reference(reference_modules):
    RP = reference_modules["partition"]
    import random
    random.seed(1)
    for _ in range(1500):
        n = random.randint(0, 40)
        iv = sorted([(a, a + random.choice([random.randint(1, 60), random.randint(100, 400)]))
                     for a in [random.randint(0, 800) for _ in range(n)]], key=lambda t: t[0])
        assert RP.group_intervals(list(iv), 128) == O.group_intervals(list(iv), 128)
    for W, seed in [(300, 1), (1000, 2), (2000, 4)]:
        m = ink_mask(synth_line(W, seed))
        ia, la, na = RP.get_binarized_islands(m, 2); ib, lb, nb = O.get_binarized_islands(m, 2)
        assert na == nb and np.array_equal(la, lb)
        assert all(np.array_equal(x[0], y[0]) and x[1] == y[1] for x, y in zip(ia, ib))
        ga = RP.group_islands(ia, (128, 128)); gb = O.group_islands(ib, (128, 128))
        assert all(np.array_equal(x[0], y[0]) and tuple(x[1]) == tuple(y[1]) for x, y in zip(ga, gb))


def test_live_sessions_vs_reference(reference_modules, parity_state):
    RB, RE = reference_modules["binarize"], reference_modules["strokes"]
    ort = O.TorchOrtSession(parity_state)
    line = synth_line(700, 12)
    a = RB.BinarizationSession().binarize_image(line, ort)
    b = O.BinarizationSessionRef().binarize_image(line, ort)
    assert np.array_equal(a, b)
    m = a[:, :, 0] > 127.5
    pa = RE.StrokeEstimationSession().get_partitions(m); pb = O.get_partitions(m)
    assert len(pa) == len(pb)
    for x, y in zip(pa, pb):
        assert np.array_equal(x["image_input"], y["image_input"]) and x["ratio"] == y["ratio"]
        assert tuple(x["translate1"]) == tuple(y["translate1"]) and tuple(x["translate2"]) == tuple(y["translate2"])


def test_resize_restatement_matches_cv2():
    """The fixed-point bilinear restatement the GPU crop kernel follows == cv2.resize (INTER_LINEAR, uint8)."""
    import cv2
    rng = np.random.default_rng(0)
    cases = [((100, 444), (222, 50)), ((1, 1), (222, 222)), ((128, 3), (5, 222)), ((2, 300), (222, 1)), ((64, 64), (32, 32))]
    for t in range(250):
        h, w = int(rng.integers(1, 140)), int(rng.integers(1, 700))
        scale = min(222 / h, 222 / w)
        cases.append(((h, w), (int(min(np.rint(scale * w), 222)), int(min(np.rint(scale * h), 222)))))
    for k, ((h, w), (dw, dh)) in enumerate(cases):
        if dw < 1 or dh < 1:
            continue
        img = (rng.random((h, w)) < (0.3 if k % 2 else 0.7)).astype(np.uint8) * 255 if k % 3 else rng.integers(0, 256, (h, w), dtype=np.uint8)
        assert np.array_equal(O.resize_linear_u8(img, dw, dh), cv2.resize(img, (dw, dh))), ((h, w), (dw, dh))


def test_resize_restatement_matches_cv2_on_rgb_lines():
    """resize_to_height (common.py:85-93) of (h, w, 3) lines: the restatement, applied per channel, == cv2.resize
    for heights below, above and at twice 128 (cv2 switches to its area path at exactly 2x)."""
    import cv2
    rng = np.random.default_rng(5)
    shapes = [(200, 1000), (64, 300), (256, 1024), (256, 1025), (130, 777), (127, 500), (1, 40), (2, 9), (129, 1290),
              (37, 400), (255, 1001), (512, 3000), (300, 2000)]
    shapes += [(int(rng.integers(3, 400)), int(rng.integers(8, 1500))) for _ in range(25)]
    for h, w in shapes:
        dw = int(w * (128 / h))
        if dw < 1:
            continue
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = O.resize_to_height(img, 128)
        mine = np.stack([O.resize_linear_u8(img[:, :, c], dw, 128) for c in range(3)], -1)
        assert ref.shape == mine.shape and np.array_equal(ref, mine), (h, w)


def test_normalize_restatement_matches_cv2():
    import cv2
    rng = np.random.default_rng(1)
    for t in range(500):
        lo = int(rng.integers(0, 200)); hi = int(rng.integers(lo, 256))
        img = rng.integers(lo, hi + 1, (int(rng.integers(1, 40)), int(rng.integers(1, 60))), dtype=np.uint8)
        assert np.array_equal(O.normalize_minmax_u8(img), cv2.normalize(img, None, 0, 255, norm_type=cv2.NORM_MINMAX))
