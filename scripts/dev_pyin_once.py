import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_cut_b200 import ops
n = int(float(sys.argv[1]) * 44100)
t = torch.arange(n, device="cuda", dtype=torch.float32) / 44100
x = 0.3 * torch.sin(2 * torch.pi * 196.0 * t) * (torch.sin(2 * torch.pi * 0.4 * t) > -0.2) + 0.02 * torch.randn(n, device="cuda")
f0, fl, vp = ops.pyin(x); torch.cuda.synchronize()
print(float(fl.float().mean()))
