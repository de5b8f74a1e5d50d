"""Generates the golden fixtures in this directory by EXECUTING THE REFERENCE
(/root/reference, unmodified, imported in place; nothing is copied).

Run in the build container only:  python tests/golden/make_golden.py
The reference modules that `import onnxruntime` are loaded with the oracle's
torch-CPU stand-in (onnxruntime is not installable offline), so the
UNet fixture is "reference Python + torch fp32 graph", not real onnxruntime.
"""

import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.dont_write_bytecode = True
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")

from oracle import segmentation_ref as O  # noqa: E402
from stroke_derenderer_b200.synth import ink_mask, synth_dense_mask, synth_line  # noqa: E402
from stroke_derenderer_b200.weights import make_parity_weights  # noqa: E402

O.install_onnxruntime_shim()
from derenderer.evaluate_binarize import BinarizationSession  # noqa: E402
from derenderer.evaluate_strokes import StrokeEstimationSession  # noqa: E402
from derenderer.helper import partition as RP  # noqa: E402
from derenderer.helper import split as RS  # noqa: E402

OUT = Path(__file__).resolve().parent


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    gold = {}
    # ---- A.1 tile geometry + cut/glue (helper/split.py) -------------------------
    geo = {}
    for W in [1, 2, 100, 383, 384, 385, 639, 640, 1000, 1536, 3072, 3840, 6144, 16384, 20480, 21000]:
        img = np.random.default_rng(W).integers(0, 256, (128, W, 3), dtype=np.uint8)
        stack, idx, widths, iw = RS.cut_and_stack([img], (1, 3, 128, 384), 64)
        out = (np.random.default_rng(W + 1).random((stack.shape[0], 1, 128, 384)) < 0.3).astype(np.uint8) * 255
        glued = RS.reconstruct_images(out, iw, idx, widths, 64)[0]
        geo[str(W)] = {"n": int(stack.shape[0]), "widths": [int(w) for w in widths[0]], "stack_sha": sha(stack),
                       "glue_sha": sha(glued), "glue_sum": int(glued.sum(dtype=np.int64))}
    gold["geometry"] = geo
    # multi-image bookkeeping (SURVEY.md A.1 example) incl. a resized image
    imgs = [np.random.default_rng(10 + i).integers(0, 256, (h, w, 3), dtype=np.uint8)
            for i, (h, w) in enumerate([(128, 300), (128, 1000), (128, 384), (200, 1000)])]
    stack, idx, widths, iw = RS.cut_and_stack(imgs, (1, 3, 128, 384), 64)
    gold["multi"] = {"shape": list(stack.shape), "indices": idx, "widths": [[int(x) for x in w] for w in widths],
                     "img_widths": [int(x) for x in iw], "stack_sha": sha(stack)}

    # ---- group_intervals known answers (helper/partition.py:248-358) -------------
    cases = [
        [(0, 300), (10, 50), (60, 100), (280, 320), (310, 330), (400, 420)],
        [(0, 100), (90, 120), (90, 140)],
        [(0, 100), (90, 140), (90, 120)],
        [],
        [(5, 9)],
        [(0, 500), (0, 500), (10, 20), (30, 600), (40, 50), (700, 900), (710, 720)],
    ]
    rng = np.random.default_rng(7)
    for _ in range(40):
        n = int(rng.integers(1, 60))
        a = np.sort(rng.integers(0, 1500, n))
        w = np.where(rng.random(n) < 0.15, rng.integers(129, 500, n), rng.integers(1, 90, n))
        cases.append([(int(x), int(x + y)) for x, y in zip(a, w)])
    gold["group_intervals"] = [{"intervals": c, "groups": RP.group_intervals(list(c), 128)} for c in cases]

    # ---- islands / groups / partitions (helper/partition.py, evaluate_strokes.py:186-224)
    se = StrokeEstimationSession()
    isl = {}
    arrays = {}
    masks = {"line300": ink_mask(synth_line(300, 1)), "line1000": ink_mask(synth_line(1000, 2)),
             "line3072": ink_mask(synth_line(3072, 0)), "dense2048": synth_dense_mask(2048, 0.01, 3),
             "empty64": np.zeros((128, 64), np.uint8), "full40": np.ones((128, 40), np.uint8)}
    big = np.zeros((128, 700), np.uint8); big[60:64, 5:690] = 1; big[10:20, 100:110] = 1; big[100:110, 300:340] = 1
    big[30:40, 650:699] = 1
    masks["long_island"] = big
    for name, m in masks.items():
        islands, labels, num = RP.get_binarized_islands(m, 2)
        groups = RP.group_islands(islands, (128, 128)) if islands else []
        parts = se.get_partitions(m) if islands else []
        arrays[f"{name}_mask"] = np.packbits(m)
        arrays[f"{name}_labels"] = labels.astype(np.int32)
        isl[name] = {
            "shape": list(m.shape), "num": int(num),
            "islands": [{"pos": [int(p[0]), int(p[1])], "shape": list(c.shape), "sha": sha(c)} for c, p in islands],
            "groups": [{"pos": [int(p[0]), int(p[1])], "shape": list(c.shape), "sha": sha(c)} for c, p in groups],
            "partitions": [{"translate1": [int(p["translate1"][0]), int(p["translate1"][1])], "ratio": float(p["ratio"]),
                            "translate2": [float(p["translate2"][0]), float(p["translate2"][1])],
                            "image_sha": sha(p["image"]), "input_sha": sha(p["image_input"])} for p in parts],
        }
    gold["islands"] = isl

    # ---- UNet through the reference's BinarizationSession (shim) ------------------
    state = make_parity_weights(123)
    ort = O.TorchOrtSession(state)
    from oracle.attunet_torch import oracle_unet_forward
    line = synth_line(1536, 5)
    bs = BinarizationSession()
    stack, idx, widths, iw = bs.preprocess_images([line])
    logits = oracle_unet_forward(ort.net, (stack / 255.).astype(np.float32), logits=True)
    shift = float(np.quantile(logits, 0.95))         # ~5 % foreground (5-10 %; DESIGN.md "precision")
    state["Conv_1x1.bias"] = (state["Conv_1x1.bias"] - shift).astype(np.float32)
    gold["unet"] = {"weights_seed": 123, "head_bias": float(state["Conv_1x1.bias"][0]), "calib": "q0.95 of logits on synth_line(1536, 5)"}
    ort = O.TorchOrtSession(state)
    x1 = np.random.default_rng(0).random((1, 3, 128, 384), dtype=np.float32)      # BASELINE config 1
    p1 = ort.run(None, {"input": x1})[0]
    arrays["config1_prob"] = p1.astype(np.float32)
    line2 = synth_line(1000, 9)
    mask = bs.binarize_image(line2, ort)
    arrays["line1000_s9_binarized"] = np.packbits(mask[:, :, 0] > 127)
    gold["unet"]["config1_prob_mean"] = float(p1.mean())
    gold["unet"]["line1000_fg"] = float((mask > 127).mean())

    (OUT / "golden.json").write_text(json.dumps(gold, separators=(",", ":")))
    np.savez_compressed(OUT / "golden_arrays.npz", **arrays)
    print("wrote", OUT / "golden.json", (OUT / "golden.json").stat().st_size, (OUT / "golden_arrays.npz").stat().st_size)


if __name__ == "__main__":
    main()
