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


def main_sharded(img_filepaths, models_folder, output_folder, strokes=True):
    """The same job over ALL GPUs of the box: `torchrun --nnodes=1 --nproc-per-node N main.py -models DIR ...`.
    One process per GPU, lines dealt to ranks by tile count, every rank's results land in its region of one shared-memory
    arena (no collective on the data path; the process group only provides the barrier), rank 0 writes every output
    file in input order — what the reference's single process does (main.py:91-136)."""
    import os
    import torch
    import torch.distributed as dist
    from stroke_derenderer_b200 import segment as _seg
    from stroke_derenderer_b200.pipeline import ShardedSegmentation
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0")) % max(torch.cuda.device_count(), 1)
    dist.init_process_group("gloo")
    paths = sorted(img_filepaths)
    # sizes of all lines: every rank decodes a slice of the folder, the (h, w) pairs are exchanged as small objects
    mine_hw = {i: load_image(paths[i]).shape[:2] for i in range(rank, len(paths), world)}
    all_hw = [None] * world
    dist.all_gather_object(all_hw, mine_hw)
    hw = {k: v for d in all_hw for k, v in d.items()}
    widths = [_seg.resized_width_hw(int(hw[i][0]), int(hw[i][1])) for i in range(len(paths))]
    cfg_b = Path(models_folder) / "configs_binarizer.json"
    bs = BinarizationSession(configs_path=str(cfg_b) if cfg_b.exists() else None, device=local)
    weights = Path(models_folder) / "binarizer.onnx"
    if not weights.exists():
        weights = Path(models_folder) / "binarizer.npz"
    ort_bs = bs.init_onnx_inference(str(weights))
    job = ShardedSegmentation(ort_bs, [load_image(paths[i]) for i in ShardedSegmentation.shard_of(widths, rank, world)], widths,
                              rank=rank, world=world, barrier=dist.barrier, lines_per_chunk=bs.lines_per_chunk, crops=strokes)
    try:
        _, got = job.step()
        if rank == 0:
            Path(output_folder).mkdir(parents=True, exist_ok=True)
            for i, p in enumerate(paths):
                stem = Path(p).stem
                img_bin = got.mask(i) > (255 * bs.bin_thr)                  # main.py:108
                save_image(normalize_image(img_bin.astype(np.uint8)), str(Path(output_folder) / f"{stem}_BINARIZED.png"), grayscale=True)
                if strokes:
                    g = got.groups(i)
                    _, ratio, t2 = _seg.crop_geometry(g) if len(g) else (None, [], [])
                    out = [{"translate1": [int(g[k, 1]), int(g[k, 2])], "ratio": float(ratio[k]),
                            "translate2": [float(t2[k, 0]), float(t2[k, 1])]} for k in range(len(g))]
                    save_json(out, str(Path(output_folder) / f"{stem}_PARTITIONS.json"))
            print(f"{len(paths)} images binarized" + (" and partitioned" if strokes else "") + f" on {world} GPU process(es); results in {output_folder}")
            del got
        job.release()
    finally:
        job.close()
        ort_bs.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    import os
    vargs = parse_args()
    paths = [str(x) for x in Path(vargs.input).glob("*.png")]
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        main_sharded(paths, vargs.models, vargs.output, strokes=True)
    else:
        sessions = initialize_sessions(vargs.models)
        main(load_images(paths), *sessions, vargs.output, strokes=True)
