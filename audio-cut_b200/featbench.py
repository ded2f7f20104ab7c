"""Feature-only measurement (BASELINE.json configs[2]): every framewise series of SURVEY.md section 8 rows
A9-A11, A17-A19 over a long synthetic mono signal that is already resident in HBM, timed per C-ABI call
with CUDA events.  ``achieved`` = algorithmic bytes (4 B per input sample + 4 B per output value) / time,
to be read against the measured HBM copy bandwidth; the STFT- and YIN-based series are compute (FFT /
autocorrelation) bound, which the GFLOP/s column makes visible."""
from __future__ import annotations

from typing import Dict

import torch

from . import ops
from .gpu_pipeline import chunk_schedule

SR = 44100


def _time(fn, iters=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def feature_only_bench(seconds: float = 3600.0, device: int = 0, hbm_peak_gbs: float = 6468.0) -> Dict:
    dev = torch.device("cuda", device)
    n = int(seconds * SR)
    g = torch.Generator(device=dev).manual_seed(7)
    t = torch.arange(n, device=dev, dtype=torch.float32) / SR
    x = 0.3 * torch.sin(2 * torch.pi * 196.0 * t) * (torch.sin(2 * torch.pi * 0.4 * t) > -0.2) \
        + 0.02 * torch.randn(n, device=dev, generator=g)
    del t
    rows = []

    def add(name, ms, out_values, flops=0.0):
        b = 4.0 * n + 4.0 * out_values
        rows.append({"kernel": name, "ms": round(ms, 3), "algorithmic_GB": round(b / 1e9, 3), "GBps": round(b / ms / 1e6, 1),
                     "hbm_frac": round(b / ms / 1e6 / hbm_peak_gbs, 3), "GFLOPps": round(flops / ms / 1e6, 1) if flops else None,
                     "x_realtime": round(seconds / (ms / 1e3), 0)})

    for frame, hop in ((4410, 2205), (1102, 441), (2048, 441), (2205, 882)):
        nf = ops.frame_count(n, frame, hop)
        out = torch.empty(nf, device=dev)
        add(f"frame_rms {frame}/{hop}", _time(lambda: ops.frame_rms(x, frame, hop, out=out)), nf)
    nf441 = 1 + n // 441
    add("zero_crossing_rate 2048/441", _time(lambda: ops.zero_crossing_rate(x, 2048, 441)), nf441)
    fft_flops = lambda frames: frames * 0.5 * 5.0 * 2048 * 11  # two real frames per complex 2048-point FFT
    add("stft2048 flatness+centroid+low_ratio @441",
        _time(lambda: ops.stft_features(x, [(0, n, 0)], 441, SR, total_frames=nf441, want=("flatness", "centroid", "low_ratio"))),
        3 * nf441, fft_flops(nf441))
    # TrackFeatureCache geometry: one segment per 10 s pipeline chunk (per-chunk top_db), hop 2205
    plans = chunk_schedule(seconds, chunk_s=10.0, overlap_s=2.5, halo_s=0.5)
    segs, off = [], 0
    for p in plans:
        cs, ce, _, _ = p.sample_bounds(SR, n)
        segs.append((cs, ce - cs, off))
        off += 1 + (ce - cs) // 2205
    add(f"stft2048 flatness+onset @2205, {len(segs)} chunk segments",
        _time(lambda: ops.stft_features(x, segs, 2205, SR, total_frames=off, want=("flatness", "onset_mean"))), 2 * off, fft_flops(off))
    nf512 = 1 + n // 512
    add("stft2048 onset mean+median @512 (BPM envelope)",
        _time(lambda: ops.stft_features(x, [(0, n, 0)], 512, SR, total_frames=nf512, want=("onset_mean", "onset_median"))),
        2 * nf512, fft_flops(nf512))
    add("pyin: YIN + candidates @441", _time(lambda: ops.pyin(x, SR, 441, decode=False), 1), nf441, 2.0 * nf441 * 676 * 1024 * 2)
    add("pyin: full (YIN + fp64 Viterbi + backtrack)", _time(lambda: ops.pyin(x, SR, 441), 1), 3 * nf441)
    add("lpc formants 1102/441 (Burg 12)", _time(lambda: ops.lpc_formants(x, SR, 441, 12), 1), 4 * (n // 441), 8.0 * (n // 441) * 12 * 1102)
    total_ms = sum(r["ms"] for r in rows)
    return {"workload": f"feature-only, {seconds:.0f} s synthetic mono, resident in HBM (configs[2])", "n_samples": n,
            "hbm_peak_GBps": hbm_peak_gbs, "total_ms": round(total_ms, 2), "x_realtime_all_series": round(seconds / (total_ms / 1e3), 0),
            "kernels": rows}
