"""CPU emulation of where the bf16 tensor-core path rounds, to find which rounding points cost the stem SDR.

Runs the oracle TFC-TDF net (fp32) on the spectrogram of a real synthetic-track window, then variants with
bf16 rounding switched on at: the input spectrogram, the weights, every stored activation, the residual
sum, the skip product, the output spectrogram.  Dev tool (uses oracle/, never imported by the package).

    python scripts/dev_bf16_emulate.py [n_windows]
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from audio_cut_b200 import synth, unet_weights as uw  # noqa: E402
from oracle import mdx  # noqa: E402
from oracle import unet as ounet  # noqa: E402


def r16(x):
    return x.bfloat16().float()


class Emu:
    def __init__(self, net, *, w16=True, in16=True, act16=True, res16=True, skip16=True, out16=True, tdf_h16=True,
                 split_in=False, levels=None):
        self.net, self.w16, self.in16, self.act16, self.res16, self.skip16, self.out16 = net, w16, in16, act16, res16, skip16, out16
        self.tdf_h16 = tdf_h16
        self.split_in = split_in
        self.levels = levels  # restrict activation rounding to these encoder depths (None = all)

    def W(self, w):
        return r16(w) if self.w16 else w

    def A(self, x, on=True):
        return r16(x) if (self.act16 and on) else x

    def conv_bn_relu(self, seq, x, fn):
        conv, bn = seq[0], seq[1]
        y = fn(x, self.W(conv.weight), None)
        y = y + conv.bias[None, :, None, None]
        y = torch.relu(nn.functional.batch_norm(y, bn.running_mean, bn.running_var, bn.weight, bn.bias, False, 0.0, bn.eps))
        return y

    def block(self, blk, x):
        for h in blk.tfc.H:
            x = self.A(self.conv_bn_relu(h, x, lambda a, w, b: nn.functional.conv2d(a, w, b, 1, 1)))
        lin1, bn1, _, lin2, bn2, _ = blk.tdf
        h1 = torch.relu(nn.functional.batch_norm(x @ self.W(lin1.weight).t(), bn1.running_mean, bn1.running_var, bn1.weight, bn1.bias, False, 0.0, bn1.eps))
        if self.tdf_h16:
            h1 = self.A(h1)
        h2 = torch.relu(nn.functional.batch_norm(h1 @ self.W(lin2.weight).t(), bn2.running_mean, bn2.running_var, bn2.weight, bn2.bias, False, 0.0, bn2.eps))
        y = x + h2
        return r16(y) if (self.res16 and self.act16) else y

    def __call__(self, x):
        n = self.net
        with torch.no_grad():
            if self.split_in:  # hi + lo bf16 split of the input: 8 input channels, weights duplicated
                hi = r16(x)
                lo = r16(x - hi)
                conv, bn = n.first_conv[0], n.first_conv[1]
                w = self.W(conv.weight)
                y = nn.functional.conv2d(hi, w) + nn.functional.conv2d(lo, w) + conv.bias[None, :, None, None]
                x = torch.relu(nn.functional.batch_norm(y, bn.running_mean, bn.running_var, bn.weight, bn.bias, False, 0.0, bn.eps))
                x = self.A(x)
            else:
                if self.in16:
                    x = r16(x)
                x = self.A(self.conv_bn_relu(n.first_conv, x, lambda a, w, b: nn.functional.conv2d(a, w, b)))
            x = x.transpose(-1, -2)
            skips = []
            for i in range(n.n):
                x = self.block(n.encoding_blocks[i], x)
                skips.append(x)
                x = self.A(self.conv_bn_relu(n.ds[i], x, lambda a, w, b: nn.functional.conv2d(a, w, b, 2)))
            x = self.block(n.bottleneck_block, x)
            for i in range(n.n):
                x = self.conv_bn_relu(n.us[i], x, lambda a, w, b: nn.functional.conv_transpose2d(a, w, b, 2))
                x = x * skips[-i - 1]
                if self.skip16:
                    x = self.A(x)
                x = self.block(n.decoding_blocks[i], x)
            x = x.transpose(-1, -2)
            fc = n.final_conv[0]
            y = nn.functional.conv2d(x, self.W(fc.weight) if False else fc.weight, fc.bias)
            return r16(y) if self.out16 else y


def sdr(ref, est):
    ref, est = ref.double(), est.double()
    return float(10 * torch.log10((ref ** 2).sum() / ((ref - est) ** 2).sum()))


def main():
    torch.set_num_threads(os.cpu_count() or 8)
    nwin = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    geo = uw.UNetGeometry()
    st = uw.random_state(geo, seed=1234)
    net = ounet.build_net(st, geo.dim_f, geo.dim_t, geo.g)
    mg = mdx.MdxGeometry(7680, 1024, 3072, 256)
    audio = synth.synth_track(12.0, seed=0, stereo=True)
    wins = mdx.build_windows(audio[:, : 10 * 44100 + 3], mg)[:nwin]
    spec = mdx.stft(torch.from_numpy(wins), mg)
    rnd = torch.randn(nwin, 4, 3072, 256, generator=torch.Generator().manual_seed(0)) * 3.0
    for tag, x in (("audio", spec), ("randn", rnd)):
        with torch.no_grad():
            ref = net(x)
        print(tag, "in rms", float(x.pow(2).mean().sqrt()), "in max", float(x.abs().max()), "out rms", float(ref.pow(2).mean().sqrt()))
        # time-domain view for the audio case: stem SDR through the iSTFT
        variants = {
            "all16": dict(),
            "in32": dict(in16=False),
            "split_in": dict(split_in=True),
            "w32": dict(w16=False),
            "act32(w16,in16,out16)": dict(act16=False),
            "only_in16": dict(w16=False, act16=False, out16=False),
            "only_w16": dict(in16=False, act16=False, out16=False),
            "only_out16": dict(in16=False, w16=False, act16=False),
            "only_act16": dict(in16=False, w16=False, out16=False),
            "act16_nores": dict(in16=False, w16=False, out16=False, res16=False),
            "act16_noskip": dict(in16=False, w16=False, out16=False, skip16=False),
        }
        for name, kw in variants.items():
            got = Emu(net, **kw)(x)
            line = f"  {tag:6s} {name:24s} spec SDR {sdr(ref, got):6.2f} dB"
            if tag == "audio":
                wr = mdx.istft(ref, mg)[:, :, mg.trim:-mg.trim].mean(1)
                wg = mdx.istft(got, mg)[:, :, mg.trim:-mg.trim].mean(1)
                line += f"   stem SDR {sdr(wr, wg):6.2f} dB"
            print(line, flush=True)


if __name__ == "__main__":
    main()
