"""Dev helper: A/B the bf16 tensor-core U-Net variants (debug mode 2 = streaming conv only, 0 = production)
at full geometry: SDR between them, time per forward, per-kernel-class profile."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_cut_b200 import ops, unet_weights as uw, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
geo = uw.UNetGeometry()
net = ops.UNet(uw.random_state(geo), geo)
x = (torch.randn(B, 256, 3072, 4, device="cuda") * 3).bfloat16()

def sdr(a, b):
    a = a.double(); b = b.double()
    return float(10 * torch.log10((a * a).sum() / ((a - b) ** 2).sum()))

def run(mode, n=3):
    net.set_debug(mode)
    out = net.forward(x); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): out = net.forward(x)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / n
    _lib.profile_begin()
    net.forward(x); torch.cuda.synchronize()
    stats = _lib.profile_collect()
    return out, t, stats

o_old, t_old, s_old = run(2)
o_new, t_new, s_new = run(0)
print(f"streaming conv : {t_old:.2f} ms / {B} windows -> {B*758.9/t_old:.1f} TFLOP/s")
print(f"production     : {t_new:.2f} ms / {B} windows -> {B*758.9/t_new:.1f} TFLOP/s")
print(f"SDR production vs streaming: {sdr(o_old.float(), o_new.float()):.1f} dB; aborted={_lib.load().ac_debug_tc_aborted()}")
for name, st in (("streaming", s_old), ("production", s_new)):
    print(name)
    for k in st:
        print("   ", k)
