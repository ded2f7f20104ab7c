"""Dev helper: pYIN on 120 s of synthetic vocal-like audio; writes f0 / flags / probabilities to the given .npz and prints the
Viterbi time (run once plain and once with AC_PYIN_GENERIC=1, then compare the files: the two paths must agree bit for bit)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audio_cut_b200 import ops, synth
y = synth.synth_track(120.0, seed=5, stereo=False).astype(np.float32)
x = torch.from_numpy(y).cuda()
f0, fl, vp = ops.pyin(x)
torch.cuda.synchronize()
t0 = time.perf_counter(); f0, fl, vp = ops.pyin(x); torch.cuda.synchronize(); t1 = time.perf_counter()
np.savez(sys.argv[1], f0=f0.cpu().numpy(), fl=fl.cpu().numpy(), vp=vp.cpu().numpy())
print(f"pyin 120 s: {1e3 * (t1 - t0):.1f} ms, voiced frames {int(fl.sum())} / {fl.numel()}")
