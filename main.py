"""Drop-in for /root/reference/main.py (CLI, initialize_sessions, load_images, main) with the
text-segmentation path on the B200.  Stroke estimation (a second model family) is outside this
framework's hot path: `strokes=True` writes the segmentation result (character
group boxes) as `<stem>_PARTITIONS.json` instead of `<stem>_STROKES.json`.

  python main.py -models DIR [-input DIR] [--output DIR]      (reference flags, main.py:20-30)
DIR holds `binarizer.onnx` (read without onnx/onnxruntime) or `binarizer.npz` (state dict with the upstream names)
and optionally `configs_binarizer.json` / `configs_strokes.json`.
"""

import argparse
import time
from pathlib import Path

import numpy as np

from stroke_derenderer_b200.common import load_image, normalize_image, save_image, save_json
from stroke_derenderer_b200.evaluate_binarize import BinarizationSession
from stroke_derenderer_b200.evaluate_strokes import StrokeEstimationSession


def parse_args(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("-models", "--models", required=True, help="Path to the folder containing all model files.")
    parser.add_argument("-input", "--input", default="./images/input", help="Folder containing all input images.")
    parser.add_argument("-output", "--output", default="./images/output", help="Output directory.")
    return parser.parse_args(argv)


def initialize_sessions(folderpath):
    """main.py:33-64 -> (bs, ort_bs, se, orts_se); orts_se is None (no stroke graphs here)."""
    folder = Path(folderpath)
    cfg_b = folder / "configs_binarizer.json"
    bs = BinarizationSession(configs_path=str(cfg_b) if cfg_b.exists() else None)
    weights = folder / "binarizer.onnx"               # main.py:43
    if not weights.exists():
        weights = folder / "binarizer.npz"
    if not weights.exists():
        raise FileNotFoundError(f"neither binarizer.onnx nor binarizer.npz found in {folder}")
    ort_bs = bs.init_onnx_inference(str(weights))
    cfg_s = folder / "configs_strokes.json"
    se = StrokeEstimationSession(configs_path=str(cfg_s) if cfg_s.exists() else None)
    return bs, ort_bs, se, None


def load_images(img_filepaths):
    """main.py:67-78."""
    return [(load_image(p), Path(p).stem) for p in img_filepaths]


def main(imgs, bs, ort_bs, se, orts_se, output_folder, strokes=True):
    """main.py:91-136, segmentation part."""
    Path(output_folder).mkdir(parents=True, exist_ok=True)
    for img, filename in imgs:
        start = time.time()
        img_bin = bs.binarize_image(img, ort_bs)
        img_bin = img_bin[:, :, 0] > (255 * bs.bin_thr)                    # main.py:108
        t_bin = round(time.time() - start, 4)
        bin_path = str(Path(output_folder) / f"{filename}_BINARIZED.png")
        save_image(normalize_image(img_bin.astype(np.uint8)), bin_path, grayscale=True)
        print(f"{filename} took {t_bin} seconds to binarize. Result is saved to {bin_path}")
        if strokes:
            start = time.time()
            parts = se.get_partitions(img_bin)
            t_se = round(time.time() - start, 4)
            out = [{"translate1": [int(p["translate1"][0]), int(p["translate1"][1])], "ratio": float(p["ratio"]),
                    "translate2": [float(p["translate2"][0]), float(p["translate2"][1])]} for p in parts]
            path = str(Path(output_folder) / f"{filename}_PARTITIONS.json")
            save_json(out, path)
            print(f"{filename} took {t_se} seconds to partition into {len(parts)} crops. Result is saved to {path}")


if __name__ == "__main__":
    vargs = parse_args()
    paths = [str(x) for x in Path(vargs.input).glob("*.png")]
    sessions = initialize_sessions(vargs.models)
    main(load_images(paths), *sessions, vargs.output, strokes=True)
