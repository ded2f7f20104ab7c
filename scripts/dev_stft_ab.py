"""Dev helper: MDX STFT / iSTFT (16-window batch, Kim_Vocal geometry) timing and parity vs torch.stft / istft.
AC_NO_FFT3=1 selects the generic mixed-radix kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_cut_b200 import _lib, ops

n_fft = int(sys.argv[1]) if len(sys.argv) > 1 else 7680
geom = ops.mdx_geom(n_fft, 1024, 3072 if n_fft == 7680 else 2048, 256)
dim_f = 3072 if n_fft == 7680 else 2048
W = 1024 * 255
torch.manual_seed(0)
wave = torch.randn(16, 2, W, device="cuda") * 0.3

def sdr(a, b):
    a = a.double(); b = b.double()
    return float(10 * torch.log10((a * a).sum() / ((a - b) ** 2).sum().clamp_min(1e-300)))

def timed(fn, it=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it

spec32 = ops.stft_mdx(wave, geom, dtype=_lib.AC_F32)
win = torch.hann_window(n_fft, periodic=True, device="cuda")
ref = torch.stft(wave.reshape(32, W), n_fft, 1024, window=win, center=True, return_complex=True)[:, :dim_f]  # [32, F, T]
ref = torch.view_as_real(ref).reshape(16, 2, dim_f, 256, 2).permute(0, 3, 2, 1, 4).reshape(16, 256, dim_f, 4)
print("stft  fp32 vs torch.stft:", round(sdr(ref, spec32), 1), "dB")
back = ops.istft_mdx(spec32, geom)
full = torch.stft(wave.reshape(32, W), n_fft, 1024, window=win, center=True, return_complex=True)
full[:, dim_f:] = 0
ref_back = torch.istft(full, n_fft, 1024, window=win, center=True, length=W).reshape(16, 2, W)
print("istft fp32 vs torch.istft:", round(sdr(ref_back, back), 1), "dB")
spec16 = ops.stft_mdx(wave, geom, dtype=_lib.AC_F16)
print(f"stft  fp16: {timed(lambda: ops.stft_mdx(wave, geom, dtype=_lib.AC_F16))*1e3:7.1f} us   fp32: {timed(lambda: ops.stft_mdx(wave, geom, dtype=_lib.AC_F32))*1e3:7.1f} us")
print(f"istft fp16: {timed(lambda: ops.istft_mdx(spec16, geom))*1e3:7.1f} us   fp32: {timed(lambda: ops.istft_mdx(spec32, geom))*1e3:7.1f} us")
