"""Seeded synthetic test audio (there is no network access for real tracks).

The signal follows SURVEY.md section 8(d) config 1: a "vocal" made of 8 harmonic
partials with vibrato under a 2 Hz on/off phrase gate, band-limited pink noise and a
120 BPM click track, peak-normalised to 0.5; the right channel is the left one
delayed by 3 samples at 0.9 gain so that L and R are decorrelated.
"""
from __future__ import annotations

import numpy as np

SR = 44100


def synth_track(seconds: float, *, sr: int = SR, seed: int = 0, stereo: bool = True) -> np.ndarray:
    """Return float32 audio, shape (2, N) when ``stereo`` else (N,)."""
    n = int(round(seconds * sr))
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) / sr

    # vocal-like: 8 partials of a slowly gliding f0 with 5.5 Hz vibrato, gated in phrases
    f0 = 220.0 * 2.0 ** (np.sin(2 * np.pi * 0.11 * t) * 0.25)
    vib = 1.0 + 0.01 * np.sin(2 * np.pi * 5.5 * t)
    phase = 2 * np.pi * np.cumsum(f0 * vib) / sr
    voc = np.zeros(n)
    for k in range(1, 9):
        voc += np.sin(k * phase + rng.uniform(0, 2 * np.pi)) / k**1.2
    gate_raw = (np.sin(2 * np.pi * 0.5 * t + 0.3) > -0.2).astype(np.float64)  # 2 s phrases with gaps
    ramp = int(0.02 * sr)
    kern = np.hanning(2 * ramp + 1)
    kern /= kern.sum()
    gate = np.convolve(gate_raw, kern, mode="same")
    voc *= gate

    # band-limited pink-ish noise: white noise shaped by 1/sqrt(f) between 60 Hz and 12 kHz
    m = 1
    while m < n:
        m *= 2
    spec = np.fft.rfft(rng.standard_normal(m))
    freqs = np.fft.rfftfreq(m, 1.0 / sr)
    shape = np.zeros_like(freqs)
    band = (freqs >= 60.0) & (freqs <= 12000.0)
    shape[band] = 1.0 / np.sqrt(freqs[band])
    noise = np.fft.irfft(spec * shape, m)[:n]
    noise /= np.max(np.abs(noise)) + 1e-12

    # 120 BPM click track: 5 ms decaying 1 kHz bursts
    clicks = np.zeros(n)
    burst_len = int(0.03 * sr)
    bt = np.arange(burst_len) / sr
    burst = np.sin(2 * np.pi * 1000.0 * bt) * np.exp(-bt / 0.005)
    for pos in range(0, n, int(0.5 * sr)):
        seg = min(burst_len, n - pos)
        clicks[pos : pos + seg] += burst[:seg]

    mix = 0.6 * voc / (np.max(np.abs(voc)) + 1e-12) + 0.25 * noise + 0.35 * clicks
    mix *= 0.5 / (np.max(np.abs(mix)) + 1e-12)
    # -70 dB white floor (as any real recording has): keeps every STFT bin well above the float32
    # FFT rounding floor, so log-domain features (flatness, mel dB) are well conditioned
    mix = mix + 3e-4 * rng.standard_normal(n)
    mix *= 0.5 / (np.max(np.abs(mix)) + 1e-12)
    left = mix
    if not stereo:
        return left.astype(np.float32)
    right = np.zeros(n)
    right[3:] = 0.9 * left[:-3]
    return np.stack([left, right]).astype(np.float32)


def synth_song(seconds: float, *, sr: int = SR, seed: int = 0) -> np.ndarray:
    """Mono "song" with phrase structure for the cut-point chain: sung phrases of 1.5-4 s separated by
    rests of 0.3-1.4 s (the accompaniment drops out in every third rest), a quiet noise bed and a 96 BPM click track.
    float32, shape (N,), peak 0.5."""
    n = int(round(seconds * sr))
    rng = np.random.default_rng(1000 + seed)
    t = np.arange(n, dtype=np.float64) / sr
    f0 = 196.0 * 2.0 ** (np.sin(2 * np.pi * 0.07 * t + seed) * 0.3)
    phase = 2 * np.pi * np.cumsum(f0 * (1.0 + 0.008 * np.sin(2 * np.pi * 5.2 * t))) / sr
    voc = np.zeros(n)
    for k in range(1, 9):
        voc += np.sin(k * phase + rng.uniform(0, 2 * np.pi)) / k**1.1
    gate = np.zeros(n)
    acc = np.ones(n)
    pos, idx = int(0.4 * sr), 0
    while pos < n:
        ln = int(rng.uniform(1.5, 4.0) * sr)
        gate[pos : pos + ln] = rng.uniform(0.6, 1.0)
        rest = int(rng.uniform(0.3, 1.4) * sr)
        if idx % 3 == 2:
            acc[pos + ln : pos + ln + rest] = 0.05
        pos += ln + rest
        idx += 1
    ramp = int(0.015 * sr)
    kern = np.hanning(2 * ramp + 1)
    kern /= kern.sum()
    gate = np.convolve(gate, kern, mode="same")
    acc = np.convolve(acc, kern, mode="same")
    noise = rng.standard_normal(n)
    noise = np.convolve(noise, np.ones(8) / 8.0, mode="same")
    clicks = np.zeros(n)
    bl = int(0.03 * sr)
    bt = np.arange(bl) / sr
    burst = np.sin(2 * np.pi * 800.0 * bt) * np.exp(-bt / 0.006)
    for p0 in range(0, n, int(0.625 * sr)):
        seg = min(bl, n - p0)
        clicks[p0 : p0 + seg] += burst[:seg]
    mix = 0.7 * voc * gate / (np.max(np.abs(voc)) + 1e-12) + acc * (0.02 * noise / (np.max(np.abs(noise)) + 1e-12) + 0.08 * clicks)
    mix = mix + 2e-4 * rng.standard_normal(n)
    mix *= 0.5 / (np.max(np.abs(mix)) + 1e-12)
    return mix.astype(np.float32)
