"""Benchmark harness for BASELINE config 3 ("full two-stage U-Net training step with fused PDE loss"): a U-Net of the
reference's SHAPE written from scratch -- the model itself is out of the hot path's scope (dense convolutions are
cuDNN's job); only its size matters here, so that the loss is timed inside a step of
realistic weight: 4 encoder levels (64, 128, 256, 512 channels) + a 512-channel bottleneck, two 3x3 conv + activation per
block with channel dropout, 2x2 max-pooling down, 2x2 transposed convolutions up with skip concatenation, a 1x1 output convolution to one channel and
the reference's `activation_name` switch (src/unet.py:156-167, :208-214) so that the train step can move the activation
into the loss kernels.  Random initialisation; no checkpoint is loaded.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class _Block(nn.Module):
    """two 3x3 convolutions, each followed by the activation; channel dropout in between (the reference's block shape)"""

    def __init__(self, cin: int, cout: int, drop: float):
        super().__init__()
        self.a = nn.Conv2d(cin, cout, 3, padding=1)
        self.b = nn.Conv2d(cout, cout, 3, padding=1)
        self.drop = nn.Dropout2d(drop) if drop > 0 else nn.Identity()

    def forward(self, x):
        return F.relu(self.b(self.drop(F.relu(self.a(x)))))


class UNetHarness(nn.Module):
    def __init__(self, base: int = 64, activation: str = "sigmoid", dropout: float = 0.2):
        super().__init__()
        w = [base, base * 2, base * 4, base * 8]
        d = [0.0, dropout * 0.5, dropout, dropout]
        self.enc = nn.ModuleList([_Block(1 if i == 0 else w[i - 1], w[i], d[i]) for i in range(4)])
        self.mid = _Block(w[3], w[3], dropout)                       # the bottleneck keeps 8 x base channels
        ups = [(w[3], w[3]), (w[3], w[2]), (w[2], w[1]), (w[1], w[0])]
        self.up = nn.ModuleList([nn.ConvTranspose2d(ci, co, 2, stride=2) for ci, co in ups])
        self.dec = nn.ModuleList([_Block(2 * w[3], w[3], dropout), _Block(2 * w[2], w[2], dropout * 0.5),
                                  _Block(2 * w[1], w[1], dropout * 0.5), _Block(2 * w[0], w[0], 0.0)])
        self.out_conv = nn.Conv2d(w[0], 1, 1)
        self.activation_name = activation

    def features(self, x):
        """The 64-channel full-resolution map the 1x1 output convolution reads (for the fused model tail)."""
        skips = []
        for blk in self.enc:
            x = blk(x)
            skips.append(x)
            x = F.max_pool2d(x, 2)
        x = self.mid(x)
        for up, dec, skip in zip(self.up, self.dec, reversed(skips)):
            x = dec(torch.cat([up(x), skip], dim=1))
        return x

    def forward(self, x):
        z = self.out_conv(self.features(x))
        if self.activation_name == "sigmoid":
            return torch.sigmoid(z)
        if self.activation_name == "tanh":
            return (torch.tanh(z) + 1.0) / 2.0
        return z


def n_params(model: nn.Module) -> int:
    return sum(p.numel() for p in model.parameters())
