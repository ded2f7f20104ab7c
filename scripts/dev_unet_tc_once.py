"""Dev helper for ncu: one bf16 tensor-core U-Net forward (after one warm-up) at full geometry."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_cut_b200 import ops, unet_weights as uw, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
geo = uw.UNetGeometry()
net = ops.UNet(uw.random_state(geo), geo)
x = (torch.randn(B, 256, 3072, 4, device="cuda") * 3).bfloat16()
net.forward(x); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record(); net.forward(x); e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1)
print(f"bf16 tc: {t:.2f} ms / {B} windows -> {B*758.9/t:.1f} TFLOP/s; aborted={_lib.load().ac_debug_tc_aborted()}")
