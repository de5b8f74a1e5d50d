"""ORACLE (test infrastructure, never on the product path).

Torch-CPU fp32 restatement of the binarizer graph that the reference executes
through onnxruntime (`/root/reference/derenderer/evaluate_binarize.py:48-53`
creates the session, `:99-100` feeds `{"input": f32 NCHW in [0,1]}` and reads
output[0]).  The graph itself is NOT in the reference tree: README.md:54 only
names it ("UNet model with attention" from namdvt/skeletonization, i.e. the
`AttU_Net` of LeeJunHyun/Image_Segmentation, un-vendored and unpinned), and the
third-party runtime is onnxruntime==1.18 (`setup.py:30`), which is absent from
this image.  PARITY UNPINNED: there is no golden vector for this graph in the
reference; this module restates the published topology
and is the fp32 checker for the CUDA path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import this file.
"""

from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn


class conv_block(nn.Module):
    def __init__(self, ch_in, ch_out):
        super().__init__()
        self.conv = nn.Sequential(
            nn.Conv2d(ch_in, ch_out, 3, 1, 1, bias=True), nn.BatchNorm2d(ch_out), nn.ReLU(inplace=True),
            nn.Conv2d(ch_out, ch_out, 3, 1, 1, bias=True), nn.BatchNorm2d(ch_out), nn.ReLU(inplace=True))

    def forward(self, x):
        return self.conv(x)


class up_conv(nn.Module):
    def __init__(self, ch_in, ch_out):
        super().__init__()
        self.up = nn.Sequential(
            nn.Upsample(scale_factor=2),
            nn.Conv2d(ch_in, ch_out, 3, 1, 1, bias=True), nn.BatchNorm2d(ch_out), nn.ReLU(inplace=True))

    def forward(self, x):
        return self.up(x)


class Attention_block(nn.Module):
    def __init__(self, F_g, F_l, F_int):
        super().__init__()
        self.W_g = nn.Sequential(nn.Conv2d(F_g, F_int, 1, 1, 0, bias=True), nn.BatchNorm2d(F_int))
        self.W_x = nn.Sequential(nn.Conv2d(F_l, F_int, 1, 1, 0, bias=True), nn.BatchNorm2d(F_int))
        self.psi = nn.Sequential(nn.Conv2d(F_int, 1, 1, 1, 0, bias=True), nn.BatchNorm2d(1), nn.Sigmoid())
        self.relu = nn.ReLU(inplace=True)

    def forward(self, g, x):
        psi = self.psi(self.relu(self.W_g(g) + self.W_x(x)))
        return x * psi


class AttU_Net(nn.Module):
    def __init__(self, img_ch=3, output_ch=1, base=64, final_sigmoid=True):
        super().__init__()
        c = [base, base * 2, base * 4, base * 8, base * 16]
        self.Maxpool = nn.MaxPool2d(2, 2)
        self.Conv1 = conv_block(img_ch, c[0])
        self.Conv2 = conv_block(c[0], c[1])
        self.Conv3 = conv_block(c[1], c[2])
        self.Conv4 = conv_block(c[2], c[3])
        self.Conv5 = conv_block(c[3], c[4])
        self.Up5 = up_conv(c[4], c[3]); self.Att5 = Attention_block(c[3], c[3], c[3] // 2); self.Up_conv5 = conv_block(c[4], c[3])
        self.Up4 = up_conv(c[3], c[2]); self.Att4 = Attention_block(c[2], c[2], c[2] // 2); self.Up_conv4 = conv_block(c[3], c[2])
        self.Up3 = up_conv(c[2], c[1]); self.Att3 = Attention_block(c[1], c[1], c[1] // 2); self.Up_conv3 = conv_block(c[2], c[1])
        self.Up2 = up_conv(c[1], c[0]); self.Att2 = Attention_block(c[0], c[0], c[0] // 2); self.Up_conv2 = conv_block(c[1], c[0])
        self.Conv_1x1 = nn.Conv2d(c[0], output_ch, 1, 1, 0)
        self.final_sigmoid = final_sigmoid

    def forward(self, x, return_logits=False, taps=None):
        x1 = self.Conv1(x)
        x2 = self.Conv2(self.Maxpool(x1))
        x3 = self.Conv3(self.Maxpool(x2))
        x4 = self.Conv4(self.Maxpool(x3))
        x5 = self.Conv5(self.Maxpool(x4))
        d5 = self.Up5(x5); a4 = self.Att5(g=d5, x=x4); d5 = self.Up_conv5(torch.cat((a4, d5), 1))
        d4 = self.Up4(d5); a3 = self.Att4(g=d4, x=x3); d4 = self.Up_conv4(torch.cat((a3, d4), 1))
        d3 = self.Up3(d4); a2 = self.Att3(g=d3, x=x2); d3 = self.Up_conv3(torch.cat((a2, d3), 1))
        d2 = self.Up2(d3); a1 = self.Att2(g=d2, x=x1); d2 = self.Up_conv2(torch.cat((a1, d2), 1))
        logits = self.Conv_1x1(d2)
        if taps is not None:
            taps.update(x1=x1, x2=x2, x3=x3, x4=x4, x5=x5, a4=a4, a3=a3, a2=a2, a1=a1,
                        d5=d5, d4=d4, d3=d3, d2=d2, logits=logits)
        if return_logits or not self.final_sigmoid:
            return logits
        return torch.sigmoid(logits)


def build_oracle_net(state: dict[str, np.ndarray], img_ch=3, output_ch=1, base=64) -> AttU_Net:
    """fp32 eval-mode module carrying exactly the arrays in `state`."""
    net = AttU_Net(img_ch, output_ch, base)
    sd = {k: torch.from_numpy(np.asarray(v, dtype=np.float32).copy()) for k, v in state.items()}
    for k in net.state_dict():
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros((), dtype=torch.long)
    net.load_state_dict(sd, strict=True)
    net.eval()
    return net


@torch.no_grad()
def oracle_unet_forward(net: AttU_Net, x: np.ndarray, logits: bool = False, batch: int = 8) -> np.ndarray:
    """x: (B,3,H,W) f32 in [0,1] -> (B,1,H,W) f32 probabilities (or logits)."""
    outs = []
    for s in range(0, x.shape[0], batch):
        xb = torch.from_numpy(np.ascontiguousarray(x[s:s + batch], dtype=np.float32))
        outs.append(net(xb, return_logits=logits).numpy())
    if not outs:
        return np.zeros((0, 1) + tuple(x.shape[2:]), np.float32)
    return np.concatenate(outs, 0)
