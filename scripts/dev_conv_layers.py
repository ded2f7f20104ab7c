"""Dev helper: per-shape timing + parity of the 3x3 conv implementations (0 simt, 1 streaming tc, 2 ws tc)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audio_cut_b200 import ops, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
impls = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2]
shapes = [(48, 256, 3072), (96, 128, 1536), (144, 64, 768), (192, 32, 384), (240, 16, 192), (288, 8, 96)]
if len(sys.argv) > 3: shapes = shapes[:int(sys.argv[3])]
rng = np.random.default_rng(0)

def sdr(a, b):
    a = a.double(); b = b.double()
    return float(10 * torch.log10((a * a).sum() / ((a - b) ** 2).sum().clamp_min(1e-30)))

for C, T, F in shapes:
    x = (torch.randn(B, T, F, C, device="cuda")).bfloat16()
    w = (rng.standard_normal((C, C, 3, 3)) / np.sqrt(9 * C)).astype(np.float32)
    scale = torch.rand(C, device="cuda") + 0.5
    shift = torch.randn(C, device="cuda") * 0.1
    gf = 2 * 9 * B * T * F * C * C / 1e9
    gb = 4 * B * T * F * C / 1e9
    ref = None
    line = f"C={C:3d} T={T:3d} F={F:4d}  {gf:7.1f} GF {gb*1e3:7.1f} MB :"
    for impl in impls:
        try:
            y, ms = ops.debug_conv3x3(x, w, scale, shift, impl, iters=6)
        except Exception as e:
            line += f"  impl{impl}: n/a"
            continue
        if ref is None:
            ref = y
            s = float("inf")
        else:
            s = sdr(ref.float(), y.float())
        line += f"  impl{impl}: {ms*1e3:7.1f} us {gf/ms:7.1f} TF/s {gb/ms*1e3:6.0f} GB/s sdr {s:5.1f}"
    print(line, flush=True)
print("aborted:", _lib.load().ac_debug_tc_aborted())
