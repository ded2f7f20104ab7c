"""Dev helper (timing / ncu): one 16-bit tensor-core U-Net forward (after warm-up) at full geometry.

    python scripts/dev_unet_tc_once.py [batch] [fp16|bf16] [iters]      (AC_UNET_SB=a,b,c selects the per-level sub-batches)
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_cut_b200 import ops, unet_weights as uw, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
fmt = sys.argv[2] if len(sys.argv) > 2 else "fp16"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 1
geo = uw.UNetGeometry()
net = ops.UNet(uw.random_state(geo), geo)
x = (torch.randn(B, 256, 3072, 4, device="cuda") * 3).to(torch.float16 if fmt == "fp16" else torch.bfloat16)
for _ in range(2):
    y = net.forward(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(iters):
    y = net.forward(x)
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / iters
print(f"{fmt} tc SB={os.environ.get('AC_UNET_SB', 'default')}: {t:.2f} ms / {B} windows -> {B*758.9/t:.1f} TFLOP/s; aborted={_lib.load().ac_debug_tc_aborted()} checksum={float(y.float().abs().mean()):.6f}")
