"""One-off: find the per-BatchNorm gain table embedded in audio-cut_b200/unet_weights.py.

Runs the oracle network (CPU) on the STFT of the first model window of the synthetic
track and, layer by layer in execution order, rescales each BatchNorm so that its
post-ReLU output has RMS 1 (0.5 for the second TDF norm); the final 1x1 conv is scaled
so that the output spectrogram has half the RMS of the input one.  Prints the table.
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audio_cut_b200 import unet_weights as uw, synth
from oracle import unet as ounet, mdx

torch.set_num_threads(os.cpu_count())
geo = uw.UNetGeometry()
g = mdx.MdxGeometry(n_fft=7680)
st = uw.random_state(geo, seed=1234, gains=[], final_gain=1.0)
net = ounet.build_net(st, geo.dim_f, geo.dim_t, geo.g)
audio = synth.synth_track(12.0)
win = mdx.build_windows(audio[:, : 441000 // 4096 * 4096], g)[:1]
spec = mdx.stft(torch.from_numpy(win), g)
gains = []
order = uw.bn_prefixes(geo)
mods = dict(net.named_modules())
def mk(name):
    target = 0.5 if name.endswith("tdf.4") else 1.0
    def hook(m, i, o):
        s = float(torch.sqrt(torch.mean(torch.relu(o) ** 2))) / target
        gains.append((name, 1.0 / s))
        with torch.no_grad():
            m.weight.mul_(1.0 / s); m.bias.mul_(1.0 / s)
        return o / s
    return hook
for p in order:
    mods[p].register_forward_hook(mk(p))
with torch.no_grad():
    y = net(spec)
assert [n for n, _ in gains] == order
fg = 0.5 * float(spec.pow(2).mean().sqrt()) / float(y.pow(2).mean().sqrt())
print("GAINS =", [float(f"{v:.4g}") for _, v in gains])
print("FINAL_GAIN =", float(f"{fg:.4g}"))
