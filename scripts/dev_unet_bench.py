"""Dev helper: time the U-Net forward (fp32 / bf16 simt / bf16 tcgen05) and print SDR vs fp32."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audio_cut_b200 import ops, unet_weights as uw, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
geo = uw.UNetGeometry()
st = uw.random_state(geo)
net = ops.UNet(st, geo)
x = torch.randn(B, 256, 3072, 4, device="cuda") * 3
def run(inp, n=3):
    net.forward(inp); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): out = net.forward(inp)
    e1.record(); torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) / n
def sdr(a, b):
    a = a.double(); b = b.double()
    return float(10 * torch.log10((a * a).sum() / ((a - b) ** 2).sum()))
ref, t32 = run(x, 2)
print(f"fp32: {t32:.1f} ms/{B} windows -> {B*758.9/t32:.1f} TFLOP/s")
xb = x.bfloat16()
net.set_debug(True); o_simt, t_simt = run(xb, 2)
net.set_debug(False); o_tc, t_tc = run(xb, 5)
print(f"bf16 simt: {t_simt:.1f} ms  sdr vs fp32 {sdr(ref, o_simt.float()):.1f} dB")
print(f"bf16 tc  : {t_tc:.1f} ms -> {B*758.9/t_tc:.1f} TFLOP/s  sdr vs fp32 {sdr(ref, o_tc.float()):.1f} dB; tc vs simt {sdr(o_simt.float(), o_tc.float()):.1f} dB")
print("aborted:", _lib.load().ac_debug_tc_aborted())
