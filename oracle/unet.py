"""Oracle TFC-TDF U-Net (what the opaque MDX23 ONNX file computes) - test infrastructure.

The reference never defines the network: /root/reference/src/audio_cut/separation/backends.py:358
just runs ``session.run`` on ``Kim_Vocal_1.onnx``.  The architecture is restated from
the public KUIELab MDX-Net ``ConvTDFNet``: L=11 (5 down, bottleneck,
5 up), l=3 convs per TFC, growth g=48, k=3, TDF bottleneck factor bn=8, BatchNorm
(inference), ReLU, multiplicative skips, 1x1 first/final convs.

Weights come in as a plain ``{name: ndarray}`` dict whose names follow the public
module (tests build seeded random ones; a real checkpoint exported to ``.npz`` drops in).
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn as nn

BN_EPS = 1e-5


class TFC(nn.Module):
    def __init__(self, c, l, k):
        super().__init__()
        self.H = nn.ModuleList(
            [nn.Sequential(nn.Conv2d(c, c, k, 1, k // 2), nn.BatchNorm2d(c, eps=BN_EPS), nn.ReLU()) for _ in range(l)]
        )

    def forward(self, x):
        for h in self.H:
            x = h(x)
        return x


class TFC_TDF(nn.Module):
    def __init__(self, c, l, f, k, bn):
        super().__init__()
        self.tfc = TFC(c, l, k)
        self.tdf = nn.Sequential(
            nn.Linear(f, f // bn, bias=False), nn.BatchNorm2d(c, eps=BN_EPS), nn.ReLU(),
            nn.Linear(f // bn, f, bias=False), nn.BatchNorm2d(c, eps=BN_EPS), nn.ReLU(),
        )

    def forward(self, x):
        x = self.tfc(x)
        return x + self.tdf(x)


class ConvTDFNet(nn.Module):
    def __init__(self, dim_f=3072, dim_t=256, dim_c=4, g=48, L=11, l=3, k=3, bn=8, scale=2):
        super().__init__()
        self.dim_f, self.dim_t, self.dim_c, self.g, self.n = dim_f, dim_t, dim_c, g, L // 2
        self.first_conv = nn.Sequential(nn.Conv2d(dim_c, g, 1), nn.BatchNorm2d(g, eps=BN_EPS), nn.ReLU())
        f, c = dim_f, g
        self.encoding_blocks = nn.ModuleList()
        self.ds = nn.ModuleList()
        for _ in range(self.n):
            self.encoding_blocks.append(TFC_TDF(c, l, f, k, bn))
            self.ds.append(nn.Sequential(nn.Conv2d(c, c + g, scale, scale), nn.BatchNorm2d(c + g, eps=BN_EPS), nn.ReLU()))
            f //= 2
            c += g
        self.bottleneck_block = TFC_TDF(c, l, f, k, bn)
        self.decoding_blocks = nn.ModuleList()
        self.us = nn.ModuleList()
        for _ in range(self.n):
            self.us.append(nn.Sequential(nn.ConvTranspose2d(c, c - g, scale, scale), nn.BatchNorm2d(c - g, eps=BN_EPS), nn.ReLU()))
            f *= 2
            c -= g
            self.decoding_blocks.append(TFC_TDF(c, l, f, k, bn))
        self.final_conv = nn.Sequential(nn.Conv2d(c, dim_c, 1))

    def forward(self, x):  # x: [B, dim_c, dim_f, dim_t]
        x = self.first_conv(x)
        x = x.transpose(-1, -2)
        skips = []
        for i in range(self.n):
            x = self.encoding_blocks[i](x)
            skips.append(x)
            x = self.ds[i](x)
        x = self.bottleneck_block(x)
        for i in range(self.n):
            x = self.us[i](x)
            x = x * skips[-i - 1]
            x = self.decoding_blocks[i](x)
        x = x.transpose(-1, -2)
        return self.final_conv(x)


def build_net(state: Dict[str, np.ndarray], dim_f=3072, dim_t=256, g=48, dtype=torch.float32, **kw) -> ConvTDFNet:
    net = ConvTDFNet(dim_f=dim_f, dim_t=dim_t, g=g, **kw)
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in state.items()}
    missing = net.load_state_dict(sd, strict=False)
    assert not [m for m in missing.missing_keys if not m.endswith("num_batches_tracked")], missing
    return net.eval().to(dtype)
