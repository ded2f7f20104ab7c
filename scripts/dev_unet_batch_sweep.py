"""Dev helper: time per window of the production bf16 U-Net forward as a function of the window batch B
(the level-0 tensors are 75 MB per window; at small B a layer's output may still be in the 126 MB L2
when the next layer reads it)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_cut_b200 import ops, unet_weights as uw

geo = uw.UNetGeometry()
net = ops.UNet(uw.random_state(geo), geo)
for B in (1, 2, 3, 4, 6, 8, 12, 16, 24, 32):
    x = (torch.randn(B, 256, 3072, 4, device="cuda") * 3).bfloat16()
    for _ in range(2): net.forward(x)
    torch.cuda.synchronize()
    n = max(3, 48 // B)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): net.forward(x)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / n
    print(f"B={B:3d}: {t:7.2f} ms per forward, {t / B:6.3f} ms per window, {B * 758.9 / t:7.1f} TFLOP/s", flush=True)
    del x
