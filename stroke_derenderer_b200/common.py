"""I/O helpers with the reference's names and behaviour
(/root/reference/derenderer/common.py:13-102).  Kept on the host: SURVEY.md 8(b)
lists load_image / load_json as part of the drop-in surface, not of the hot path.
"""

import json
import pickle

import cv2

EPS = 1e-6


def load_image(img_filepath, grayscale=False):
    """common.py:13-24: RGB (H,W,3) u8, or gray (H,W,1)."""
    image = cv2.imread(img_filepath)
    if grayscale:
        return cv2.cvtColor(image, cv2.COLOR_BGR2GRAY)[:, :, None]
    return cv2.cvtColor(image, cv2.COLOR_BGR2RGB)


def save_image(img, save_filepath, grayscale=False):
    """common.py:27-34."""
    code = cv2.COLOR_GRAY2BGR if grayscale else cv2.COLOR_RGB2BGR
    cv2.imwrite(save_filepath, cv2.cvtColor(img, code))


def _dump(obj, path, mode, writer):
    with open(path, mode) as fh:
        writer(obj, fh)


def _read(path, mode, reader):
    with open(path, mode) as fh:
        return reader(fh)


# common.py:37-82: the pickle / YAML / JSON one-liners the sessions and scripts import by name
def save_metrics(metrics, filename): _dump(metrics, filename, "wb", pickle.dump)
def load_metrics(filename): return _read(filename, "rb", pickle.load)
def load_json(json_path): return _read(json_path, "r", json.load)
def save_json(json_dict, save_path): _dump(json_dict, save_path, "w", json.dump)


def load_yaml(filepath):
    import yaml                      # only needed by the training-side scripts
    return _read(filepath, "r", yaml.safe_load)


def resize_to_height(img, height):
    """common.py:85-93, host form (cv2) for the per-stage legacy API (`preprocess_images`, `cut_and_stack`).
    The fused path (`binarize_images`, `LineSegmentationJob`) resizes on the GPU instead: `segment.upload_lines`
    -> sd_resize_lines, bit-exact with this call."""
    h, w = img.shape[0], img.shape[1]
    return cv2.resize(img, (int(w * (height / h)), height))


def normalize_image(image):
    """common.py:96-102."""
    return cv2.normalize(image, None, 0, 255, norm_type=cv2.NORM_MINMAX)


def init_onnx_session(onnx_path, device=0, max_tiles=64):
    """common.py:105-111 creates an onnxruntime CPU session; here the handle is
    the B200 engine (only defined for the binarizer graph)."""
    from .engine import UNetEngine
    return UNetEngine(onnx_path, device=device, max_tiles=max_tiles)
