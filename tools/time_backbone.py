"""Times the frozen PSPNet feature extractor SEPARATELY from the head (BASELINE.json north_star: "the frozen PSPNet
ResNet-50/101 backbone stays as the reference's PyTorch module and is timed separately").

The reference checkout does not travel to the GPU box, so this tool builds a stand-in with the reference's layer
shapes — ``src/model/resnet.py`` (deep_base stem :110-116, Bottleneck blocks [3,4,6,3] / [3,4,23,3]),
``src/model/pspnet.py:103-127`` (layer3 dilation 2 / layer4 dilation 4 with stride 1, PPM bins [1,2,3,6] -> 4 x 512,
3x3 bottleneck 4096 -> 512 + BN + ReLU) — random init, ``eval()``, plain PyTorch eager fp32 (cuDNN), 473 x 473 input
-> [512, 60, 60]. It is NOT part of the product: the head consumes whatever the reference's ``extract_features`` emits.

    python tools/time_backbone.py [--layers 50|101] [--batch 2] [--iters 10]
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.nn as nn
import torch.nn.functional as F


class Bottleneck(nn.Module):
    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=dilation, dilation=dilation, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.downsample = downsample

    def forward(self, x):
        r = x if self.downsample is None else self.downsample(x)
        x = F.relu(self.bn1(self.conv1(x)))
        x = F.relu(self.bn2(self.conv2(x)))
        return F.relu(self.bn3(self.conv3(x)) + r)


class PSPNetFeatures(nn.Module):
    def __init__(self, layers=50, bins=(1, 2, 3, 6), bottleneck_dim=512):
        super().__init__()
        blocks = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3)}[layers]
        self.stem = nn.Sequential(
            nn.Conv2d(3, 64, 3, 2, 1, bias=False), nn.BatchNorm2d(64), nn.ReLU(True),
            nn.Conv2d(64, 64, 3, 1, 1, bias=False), nn.BatchNorm2d(64), nn.ReLU(True),
            nn.Conv2d(64, 128, 3, 1, 1, bias=False), nn.BatchNorm2d(128), nn.ReLU(True),
            nn.MaxPool2d(3, 2, 1))
        self.inplanes = 128
        self.layer1 = self._make(64, blocks[0], 1, 1)
        self.layer2 = self._make(128, blocks[1], 2, 1)
        self.layer3 = self._make(256, blocks[2], 1, 2)       # stride removed, dilation 2 (pspnet.py:103-107)
        self.layer4 = self._make(512, blocks[3], 1, 4)       # stride removed, dilation 4 (pspnet.py:108-112)
        red = 2048 // len(bins)
        self.ppm = nn.ModuleList(nn.Sequential(nn.AdaptiveAvgPool2d(b), nn.Conv2d(2048, red, 1, bias=False),
                                               nn.BatchNorm2d(red), nn.ReLU(True)) for b in bins)
        self.bottleneck = nn.Sequential(nn.Conv2d(4096, bottleneck_dim, 3, padding=1, bias=False),
                                        nn.BatchNorm2d(bottleneck_dim), nn.ReLU(True), nn.Dropout2d(0.1))

    def _make(self, planes, n, stride, dilation):
        down = None
        if stride != 1 or self.inplanes != planes * 4:
            down = nn.Sequential(nn.Conv2d(self.inplanes, planes * 4, 1, stride, bias=False), nn.BatchNorm2d(planes * 4))
        layers = [Bottleneck(self.inplanes, planes, stride, dilation, down)]
        self.inplanes = planes * 4
        layers += [Bottleneck(self.inplanes, planes, 1, dilation) for _ in range(1, n)]
        return nn.Sequential(*layers)

    def forward(self, x):
        x = self.layer4(self.layer3(self.layer2(self.layer1(self.stem(x)))))
        size = x.shape[2:]
        x = torch.cat([x] + [F.interpolate(f(x), size, mode="bilinear", align_corners=True) for f in self.ppm], 1)
        return self.bottleneck(x)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=50)
    ap.add_argument("--batch", type=int, default=2, help="images per forward: 2 = one 1-shot episode (query + support)")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--device", default="cuda:0" if torch.cuda.is_available() else "cpu")
    a = ap.parse_args()
    dev = torch.device(a.device)
    torch.backends.cudnn.benchmark = True
    net = PSPNetFeatures(a.layers).to(dev).eval()
    n_par = sum(p.numel() for p in net.parameters())
    x = torch.randn(a.batch, 3, 473, 473, device=dev)
    with torch.no_grad():
        for _ in range(3):
            y = net(x)
        assert tuple(y.shape) == (a.batch, 512, 60, 60), y.shape
        if dev.type == "cuda":
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters):
                net(x)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.iters
        else:
            t0 = time.perf_counter()
            for _ in range(a.iters):
                net(x)
            ms = (time.perf_counter() - t0) * 1e3 / a.iters
    print(json.dumps({"what": f"PSPNet-R{a.layers} feature extractor (stand-in with the reference's layer shapes, random init, "
                              "eval, PyTorch eager fp32)", "device": str(dev), "params_M": round(n_par / 1e6, 2),
                      "images_per_forward": a.batch, "ms_per_forward": round(ms, 3),
                      "ms_per_image": round(ms / a.batch, 3),
                      "episodes_per_s_1shot": round(1e3 / (ms / a.batch * 2), 1)}))


if __name__ == "__main__":
    sys.exit(main())
