"""Dev helper: the level-0 conv chain (3 x conv3x3, C = 48): three weight-stationary launches vs the fused kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audio_cut_b200 import ops, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = int(sys.argv[2]) if len(sys.argv) > 2 else 256
F = int(sys.argv[3]) if len(sys.argv) > 3 else 3072
C = 48
rng = np.random.default_rng(0)
x = torch.randn(B, T, F, C, device="cuda").half()
w = (rng.standard_normal((3, C, C, 3, 3)) / np.sqrt(9 * C) * 1.6).astype(np.float32)
scale = torch.rand(3, C, device="cuda") + 0.5
shift = torch.randn(3, C, device="cuda") * 0.1
gf = 3 * 2 * 9 * B * T * F * C * C / 1e9
ref, ms0 = ops.debug_conv3x3_chain(x, w, scale, shift, 0, iters=6)
print(f"3 launches : {ms0*1e3:8.1f} us  {gf/ms0:7.1f} TFLOP/s  aborted {_lib.load().ac_debug_tc_aborted()}", flush=True)
y, ms1 = ops.debug_conv3x3_chain(x, w, scale, shift, 1, iters=6)
ab = _lib.load().ac_debug_tc_aborted()
neq = int((ref.view(torch.int16) != y.view(torch.int16)).sum())
print(f"fused      : {ms1*1e3:8.1f} us  {gf/ms1:7.1f} TFLOP/s  aborted {ab}  differing elements {neq} of {ref.numel()}", flush=True)
if neq:
    d = (ref.float() - y.float())
    idx = (ref.view(torch.int16) != y.view(torch.int16)).nonzero()
    print("first mismatches (b,t,f,c):", idx[:8].tolist(), "max abs diff", float(d.abs().max()))
    bad_t = torch.unique(idx[:, 1])[:16].tolist(); bad_f = torch.unique(idx[:, 2])[:24].tolist()
    print("rows:", bad_t, "positions:", bad_f)
