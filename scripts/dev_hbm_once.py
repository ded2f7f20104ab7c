"""Dev helper for ncu: every HBM-class kernel of the path once (after one warm-up) on the bench workload's shapes:
MDX STFT / fused iSTFT for a 16-window batch of a 4-min stereo track, the STFT-2048 feature kernels, framewise RMS (whole
track and all chunks in one launch), YIN probabilities and LPC formants on a 60 s mono signal."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from audio_cut_b200 import _lib, ops, synth
from audio_cut_b200.gpu_pipeline import chunk_schedule

SR = 44100
fmt = sys.argv[1] if len(sys.argv) > 1 else "fp16"
dtype = {"fp16": _lib.AC_F16, "bf16": _lib.AC_BF16, "fp32": _lib.AC_F32}[fmt]
audio = synth.synth_track(240.0, seed=0, stereo=True)
n = audio.shape[-1]
plans = chunk_schedule(n / float(SR))
bounds = [p.sample_bounds(SR, n) for p in plans]
mix = torch.from_numpy(audio).cuda()
mono = mix.mean(dim=0)
geom = ops.mdx_geom(7680, 1024, 3072, 256)
W = 1024 * 255
wave = mix[:, : 16 * W].reshape(2, 16, W).permute(1, 0, 2).contiguous()
for rep in range(2):
    spec = ops.stft_mdx(wave, geom, dtype=dtype)
    back = ops.istft_mdx(spec, geom)
    segs, off = [], 0
    for cs, ce, _, _ in bounds:
        segs.append((cs, ce - cs, off))
        off += 1 + (ce - cs) // 2205
    ops.stft_features(mono, segs, 2205, SR, total_frames=off, want=("flatness", "onset_mean"))
    ops.stft_features(mono, [(0, n, 0)], 441, SR, total_frames=1 + n // 441, want=("flatness",))
    ops.stft_features(mono, [(0, n, 0)], 512, SR, total_frames=1 + n // 512, want=("onset_median",))
    for fr, hp in ((1102, 441), (2048, 441), (2205, 882)):
        ops.frame_rms(mono, fr, hp)
    rsegs, roff = [], 0
    for cs, ce, _, _ in bounds:
        rsegs.append((cs, ce - cs, roff))
        roff += ops.frame_count(ce - cs, 4410, 2205)
    ops.frame_rms_segments(mono, rsegs, 4410, 2205, total_frames=roff)
    ops.zero_crossing_rate(mono, 2048, 441)
    ops.pyin(mono[: 60 * SR], SR, 441, decode=False)
    ops.lpc_formants(mono[: 60 * SR], SR, 441, 12)
    torch.cuda.synchronize()
print("ok", float(back.abs().max()))
