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
    """main.py:91-136, segmentation part.  The reference walks the images one at a time (:104); here ALL images of
    the folder go through one pipelined job (tiles of different images share UNet passes), then the per-image
    files are written exactly as the reference names them."""
    from stroke_derenderer_b200 import segment as _seg
    from stroke_derenderer_b200.pipeline import segment_lines
    Path(output_folder).mkdir(parents=True, exist_ok=True)
    images = [img for img, _ in imgs]
    start = time.time()
    default_se = (se.margin, se.img_size, list(se.mean), list(se.std)) == (_seg.MARGIN, _seg.IMG_SIZE, _seg.IMAGENET_MEAN, _seg.IMAGENET_STD)
    if strokes and default_se and images:
        masks, parts = segment_lines(ort_bs, images, bin_thr=bs.bin_thr, lines_per_chunk=bs.lines_per_chunk, seg=bs._segmenter(ort_bs))
    else:
        masks = bs.binarize_images(images, ort_bs)
        parts = se.get_partitions_batch([m[:, :, 0] > (255 * bs.bin_thr) for m in masks]) if strokes else None
    t_all = round(time.time() - start, 4)
    print(f"{len(images)} images took {t_all} seconds to binarize" + (" and partition." if strokes else "."))
    for k, (img, filename) in enumerate(imgs):
        img_bin = masks[k][:, :, 0] > (255 * bs.bin_thr)                   # main.py:108
        bin_path = str(Path(output_folder) / f"{filename}_BINARIZED.png")
        save_image(normalize_image(img_bin.astype(np.uint8)), bin_path, grayscale=True)
        print(f"{filename}: result is saved to {bin_path}")
        if strokes:
            out = [{"translate1": [int(p["translate1"][0]), int(p["translate1"][1])], "ratio": float(p["ratio"]),
                    "translate2": [float(p["translate2"][0]), float(p["translate2"][1])]} for p in parts[k]]
            path = str(Path(output_folder) / f"{filename}_PARTITIONS.json")
            save_json(out, path)
            print(f"{filename}: {len(out)} crops, saved to {path}")


if __name__ == "__main__":
    vargs = parse_args()
    paths = [str(x) for x in Path(vargs.input).glob("*.png")]
    sessions = initialize_sessions(vargs.models)
    main(load_images(paths), *sessions, vargs.output, strokes=True)
