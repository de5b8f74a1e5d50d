"""Encoder handle for the stroke-estimator front end.

The reference runs `orts["encoder"].run(["output"], {"input": images})` on a Drive-hosted `encoder.onnx`
(/root/reference/derenderer/evaluate_strokes.py:150-160, :256) whose topology is named nowhere in the tree.  What the
code around it fixes: the input is (B, 3, 224, 224) ImageNet-normalised f32 (:58-69), the output is (B, C, 7, 7) — it
fills the 14 x 14 grid of `_encode_postprocess` after the 2 x 2 repeat that "replaces the AdaptiveAvgPool2d layer in
the encoder model" (:72-91) — and the decoder takes `mean(enc, axis=1)` into `decoder_init` (:264-266).  That is the
encoder of the "Show, Attend and Tell" layout: a stride-32 ResNet trunk without its pooling / fc head.  [recalled, NOT
verifiable offline: depth and width are parameters here; PARITY UNPINNED like the UNet]

This handle is the LIBRARY path for that model: a plain torch module (cuDNN convolutions, fp16 channels-last on the
GPU), with the onnxruntime call signature the reference uses plus `run_device` so the batch never leaves HBM.  It is not
one of the hand-written sm_100a kernels of the segmentation hot path; it exists so that `encode_partitions_batch` has
a real encoder to batch for (seeded weights, or a state dict with torchvision's ResNet parameter names).
"""

from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

_LAYERS = {18: ("basic", [2, 2, 2, 2]), 34: ("basic", [3, 4, 6, 3]), 50: ("bottleneck", [3, 4, 6, 3]),
           101: ("bottleneck", [3, 4, 23, 3])}


class _Basic(nn.Module):
    expansion = 1

    def __init__(self, cin, planes, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, planes, 3, stride, 1, bias=False); self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False); self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = None
        if stride != 1 or cin != planes:
            self.downsample = nn.Sequential(nn.Conv2d(cin, planes, 1, stride, bias=False), nn.BatchNorm2d(planes))

    def forward(self, x):
        y = torch.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        return torch.relu(y + (x if self.downsample is None else self.downsample(x)))


class _Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, cin, planes, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, planes, 1, bias=False); self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False); self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False); self.bn3 = nn.BatchNorm2d(planes * 4)
        self.downsample = None
        if stride != 1 or cin != planes * 4:
            self.downsample = nn.Sequential(nn.Conv2d(cin, planes * 4, 1, stride, bias=False), nn.BatchNorm2d(planes * 4))

    def forward(self, x):
        y = torch.relu(self.bn1(self.conv1(x)))
        y = torch.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        return torch.relu(y + (x if self.downsample is None else self.downsample(x)))


class ResNetTrunk(nn.Module):
    """conv1 7x7/2 + BN + ReLU + maxpool 3x3/2 + layer1..layer4 (torchvision parameter names), no avgpool / fc:
    (B, 3, H, W) -> (B, C, H/32, W/32), C = 512 (depth 18 / 34) or 2048 (50 / 101)."""

    def __init__(self, depth: int = 101, base: int = 64):
        super().__init__()
        kind, counts = _LAYERS[depth]
        block = _Basic if kind == "basic" else _Bottleneck
        self.conv1 = nn.Conv2d(3, base, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(base)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        cin, stages = base, []
        for i, n in enumerate(counts):
            planes = base << i
            blocks = []
            for j in range(n):
                blocks.append(block(cin, planes, 2 if (j == 0 and i > 0) else 1))
                cin = planes * block.expansion
            stages.append(nn.Sequential(*blocks))
        self.layer1, self.layer2, self.layer3, self.layer4 = stages
        self.out_channels = cin

    def forward(self, x):
        x = self.maxpool(torch.relu(self.bn1(self.conv1(x))))
        return self.layer4(self.layer3(self.layer2(self.layer1(x))))


def seeded_trunk(depth: int = 101, seed: int = 7) -> ResNetTrunk:
    """A trunk with a well-conditioned seeded initialisation (He weights; BN scale / shift / running statistics drawn
    like the binarizer's parity recipe) for tests and synthetic runs."""
    net = ResNetTrunk(depth)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, nn.Conv2d):
                fan_in = m.in_channels * m.kernel_size[0] * m.kernel_size[1]
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * (2.0 / fan_in) ** 0.5)
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.copy_(0.5 + 0.5 * torch.rand(m.weight.shape, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
                m.running_mean.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
                m.running_var.copy_(0.75 + 0.5 * torch.rand(m.bias.shape, generator=g))
    return net.eval()


class EncoderHandle:
    """`orts["encoder"]` of evaluate_strokes.py: `.run(["output"], {"input": f32 (B,3,224,224)}) -> [f32 (B,C,7,7)]`, and
    `.run_device` with cuda tensors.  fp16 channels-last on the GPU (cuDNN), fp32 out."""

    def __init__(self, trunk: ResNetTrunk, device: int = 0, half: bool = True, max_batch: int = 256):
        self.device = torch.device("cuda", device)
        self.dtype = torch.float16 if half else torch.float32
        self.net = trunk.eval().to(self.device, self.dtype).to(memory_format=torch.channels_last)
        self.max_batch = max_batch
        self.out_channels = trunk.out_channels

    @torch.no_grad()
    def run_device(self, output_names, feeds):
        x = feeds["input"]
        outs = []
        for s in range(0, x.shape[0], self.max_batch):
            xb = x[s:s + self.max_batch].to(self.device, self.dtype).contiguous(memory_format=torch.channels_last)
            outs.append(self.net(xb).float().contiguous())
        return [torch.cat(outs, 0) if outs else torch.zeros((0, self.out_channels, 7, 7), device=self.device)]

    def run(self, output_names, feeds):
        x = torch.from_numpy(np.ascontiguousarray(feeds["input"], dtype=np.float32))
        return [self.run_device(output_names, {"input": x})[0].cpu().numpy()]
