import json
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"
REFERENCE = Path("/root/reference")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


@pytest.fixture(scope="session")
def golden():
    return json.loads((GOLDEN / "golden.json").read_text())


@pytest.fixture(scope="session")
def golden_arrays():
    with np.load(GOLDEN / "golden_arrays.npz") as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def parity_state(golden):
    """Seeded parity weights with the calibrated head bias recorded in the golden file."""
    from stroke_derenderer_b200.weights import make_parity_weights
    st = make_parity_weights(golden["unet"]["weights_seed"])
    st["Conv_1x1.bias"] = np.array([golden["unet"]["head_bias"]], np.float32)
    return st


@pytest.fixture(scope="session")
def oracle_net(parity_state):
    from oracle.attunet_torch import build_oracle_net
    return build_oracle_net(parity_state)


@pytest.fixture(scope="session")
def reference_modules():
    """The unmodified reference, imported in place (build container only)."""
    if not (REFERENCE / "derenderer").exists():
        pytest.skip("/root/reference not present on this machine")
    from oracle import segmentation_ref as O
    O.install_onnxruntime_shim()
    sys.dont_write_bytecode = True
    if str(REFERENCE) not in sys.path:
        sys.path.insert(0, str(REFERENCE))
    import derenderer.evaluate_binarize as RB
    import derenderer.evaluate_strokes as RE
    import derenderer.helper.partition as RP
    import derenderer.helper.split as RS
    return {"split": RS, "partition": RP, "binarize": RB, "strokes": RE}


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return 0
