"""UNet parity of ONE operand-type build against the torch-CPU fp32 oracle, as one JSON line:
    SD_DTYPE=f16  python tools/parity_report.py      (libsd_b200.so, the parity build)
    SD_DTYPE=bf16 python tools/parity_report.py      (libsd_b200_bf16.so, same sources with -DSD_BF16)
Cases: BASELINE config 1 (uniform-noise tile, committed golden probabilities), 8 text tiles, BASELINE config 2
(one 3072-px line through the fused path).  Bars of the north star: prob max-abs 2e-2, masks >= 99.9 %.
The oracle is test infrastructure (oracle/): this tool is a checker, like tests/.
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import segmentation_ref as O  # noqa: E402
from oracle.attunet_torch import build_oracle_net, oracle_unet_forward  # noqa: E402
from stroke_derenderer_b200 import _lib, segment as S  # noqa: E402
from stroke_derenderer_b200.engine import UNetEngine  # noqa: E402
from stroke_derenderer_b200.synth import synth_line  # noqa: E402
from stroke_derenderer_b200.weights import make_parity_weights  # noqa: E402


def main():
    gold = json.loads((ROOT / "tests" / "golden" / "golden.json").read_text())
    state = make_parity_weights(gold["unet"]["weights_seed"])
    state["Conv_1x1.bias"] = np.array([gold["unet"]["head_bias"]], np.float32)
    gz = np.load(ROOT / "tests" / "golden" / "golden_arrays.npz")
    out = {"dtype": _lib.lib().sd_operand_dtype().decode(), "prob_bar": 2e-2, "mask_bar": 0.999,
           "oracle": "torch-CPU fp32 of the published topology (unpinned)"}
    eng = UNetEngine(state, device=0, max_tiles=16)

    def cmp(prob, ref):
        return {"prob_max_abs": float(np.abs(prob - ref).max()), "mask_agree": float(((prob > 0.5) == (ref > 0.5)).mean()),
                "fg_ref": float((ref > 0.5).mean())}
    x1 = np.random.default_rng(0).random((1, 3, 128, 384), dtype=np.float32)
    out["config1"] = cmp(eng.run(None, {"input": x1})[0], gz["config1_prob"])
    net = build_oracle_net(state)
    line = synth_line(320 * 8 + 200, 48)
    stack, *_ = O.cut_and_stack([line], (1, 3, 128, 384), 64)
    x8 = (stack[:8] / 255.).astype(np.float32)
    out["tiles8"] = cmp(eng.run(None, {"input": x8})[0], oracle_unet_forward(net, x8))
    l2 = synth_line(3072, 0)
    batch, planes = S.Segmenter(eng).binarize([l2])
    torch.cuda.synchronize()
    mask = batch.plane(planes, 0).cpu().numpy()
    ref = O.BinarizationSessionRef().binarize_image(l2, O.TorchOrtSession(state))[:, :, 0]
    out["config2_line"] = {"mask_agree": float(((mask > 127) == (ref > 127)).mean()), "fg_ref": float((ref > 127).mean())}
    eng.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
