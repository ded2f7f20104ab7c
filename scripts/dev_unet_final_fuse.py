"""Dev helper: the network's last TDF2 with the final 1x1 conv fused into its epilogue (production, debug mode 0)
against the two separate kernels (debug mode 3): outputs must be bit-identical; time per forward."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_cut_b200 import ops, unet_weights as uw, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
geo = uw.UNetGeometry()
net = ops.UNet(uw.random_state(geo), geo)
x = (torch.randn(B, 256, 3072, 4, device="cuda") * 3).bfloat16()

def run(mode, n=5):
    net.set_debug(mode)
    out = net.forward(x).clone(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): net.forward(x)
    e1.record(); torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) / n

o3, t3 = run(3)
o0, t0 = run(0)
o3b, t3b = run(3)
o0b, t0b = run(0)
print(f"separate: {t3:.3f} / {t3b:.3f} ms   fused: {t0:.3f} / {t0b:.3f} ms per {B} windows")
print("bit-identical:", bool(torch.equal(o3, o0)), " max abs diff:", float((o3.float() - o0.float()).abs().max()),
      " aborted:", _lib.load().ac_debug_tc_aborted())
for rep in range(3):
    for mode in (3, 0):
        _, t = run(mode, n=10)
        print(f"  rep {rep} mode {mode}: {t:.3f} ms")
for mode in (3, 0):
    net.set_debug(mode)
    _lib.profile_begin(); net.forward(x); torch.cuda.synchronize()
    print("mode", mode)
    for k in _lib.profile_collect():
        if k["name"] in ("tdf_tcgen05", "conv1x1"): print("   ", k["name"], k["launches"], round(k["total_ms"], 3))
