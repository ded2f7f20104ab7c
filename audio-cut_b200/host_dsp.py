"""Host-side (numpy) sequential logic that consumes the GPU-produced series.

SURVEY.md section 8(a) keeps these on the host in v1: they are data-dependent scans over a few
thousand frames (rows A12, A14), not bandwidth work.  They restate the published librosa (>=0.10)
algorithms named at each function; librosa itself is not a dependency of this package.

  * ``onset_detect`` / ``peak_pick``   librosa.onset.onset_detect      features_cache.py:186
  * ``tempogram`` / ``tempo``          librosa.feature.rhythm.tempo    features_cache.py:283-288,
                                                                       adaptive_vad_enhancer.py:151-156
  * ``beat_track``                     librosa.beat.beat_track         features_cache.py:289-294,
                                                                       adaptive_vad_enhancer.py:61-67
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import scipy.fft
import scipy.signal


def peak_pick(x: np.ndarray, pre_max: int, post_max: int, pre_avg: int, post_avg: int, delta: float, wait: int) -> np.ndarray:
    """x[n] is a peak when it equals max(x[n-pre_max : n+post_max]), is at least
    mean(x[n-pre_avg : n+post_avg]) + delta, is non-zero, and lies more than ``wait`` after the last peak."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    pre_max, post_max, pre_avg, post_avg, wait = (int(np.ceil(v)) for v in (pre_max, post_max, pre_avg, post_avg, wait))
    idx = np.arange(n)
    # running max / mean over clipped windows via prefix structures (vectorised)
    csum = np.concatenate([[0.0], np.cumsum(x)])
    lo_a, hi_a = np.maximum(0, idx - pre_avg), np.minimum(n, idx + post_avg)
    mean = (csum[hi_a] - csum[lo_a]) / np.maximum(1, hi_a - lo_a)
    width = pre_max + post_max
    pad = np.full(pre_max, -np.inf)
    xp = np.concatenate([pad, x, np.full(max(post_max - 1, 0), -np.inf)])
    win = np.lib.stride_tricks.sliding_window_view(xp, width)[:n]
    mx = win.max(axis=1)
    cand = np.nonzero((x == mx) & (x >= mean + delta) & (x != 0))[0]
    peaks = []
    last = -np.inf
    for i in cand:
        if i > last + wait:
            peaks.append(int(i))
            last = i
    return np.asarray(peaks, dtype=np.int64)


def onset_detect(onset_envelope: np.ndarray, sr: int, hop_length: int) -> np.ndarray:
    env = np.asarray(onset_envelope)
    if env.size == 0 or not env.any() or not np.all(np.isfinite(env)):
        return np.zeros(0, dtype=np.int64)
    env = env - env.min()
    env = env / (env.max() + np.finfo(env.dtype if env.dtype.kind == "f" else np.float32).tiny)
    return peak_pick(
        env,
        pre_max=0.03 * sr // hop_length,
        post_max=0.00 * sr // hop_length + 1,
        pre_avg=0.10 * sr // hop_length,
        post_avg=0.10 * sr // hop_length + 1,
        delta=0.07,
        wait=0.03 * sr // hop_length,
    )


def tempogram(onset_envelope: np.ndarray, win_length: int, block: int = 4096) -> np.ndarray:
    """Autocorrelation tempogram, shape (win_length, n): hann-windowed frames of the envelope
    (hop 1, centred with a linear ramp to 0), autocorrelated by FFT, max-normalised per frame."""
    env = np.asarray(onset_envelope, dtype=np.float64)
    n = env.shape[0]
    half = win_length // 2
    padded = np.pad(env, (half, half), mode="linear_ramp", end_values=(0, 0))
    window = scipy.signal.get_window("hann", win_length, fftbins=True)
    nfft = scipy.fft.next_fast_len(2 * win_length - 1, real=True)
    out = np.empty((win_length, n), dtype=np.float64)
    frames = np.lib.stride_tricks.sliding_window_view(padded, win_length)
    for s in range(0, n, block):
        fr = frames[s : min(n, s + block)] * window[None, :]
        spec = scipy.fft.rfft(fr, n=nfft, axis=1)
        ac = scipy.fft.irfft(spec.real**2 + spec.imag**2, n=nfft, axis=1)[:, :win_length]
        mag = np.max(np.abs(ac), axis=1, keepdims=True)
        mag[mag < np.finfo(np.float64).tiny] = 1.0
        out[:, s : s + fr.shape[0]] = (ac / mag).T
    return out


def tempo_from_tempogram(tg: np.ndarray, sr: int, hop_length: int, start_bpm: float = 120.0, std_bpm: float = 1.0,
                         max_tempo: float = 320.0, aggregate="mean") -> np.ndarray:
    win_length = tg.shape[0]
    if aggregate == "mean":
        tg = tg.mean(axis=1, keepdims=True)
    bpms = np.empty(win_length)
    bpms[0] = np.inf
    bpms[1:] = 60.0 * sr / (hop_length * np.arange(1.0, win_length))
    with np.errstate(divide="ignore", invalid="ignore"):
        logprior = -0.5 * ((np.log2(bpms) - np.log2(start_bpm)) / std_bpm) ** 2
    max_idx = int(np.argmax(bpms < max_tempo))
    logprior[:max_idx] = -np.inf
    best = np.argmax(np.log1p(1e6 * tg) + logprior[:, None], axis=0)
    return bpms[best]


def tempo(onset_envelope: np.ndarray, sr: int, hop_length: int, aggregate="mean", **kw) -> np.ndarray:
    win = int(np.floor(8.0 * sr / hop_length))
    return tempo_from_tempogram(tempogram(onset_envelope, win), sr, hop_length, aggregate=aggregate, **kw)


def _local_score(env: np.ndarray, period: int) -> np.ndarray:
    window = np.exp(-0.5 * (np.arange(-period, period + 1) * 32.0 / period) ** 2)
    norm = env.std(ddof=1)
    return scipy.signal.convolve(env / (norm + np.finfo(env.dtype).tiny), window, "same")


def _beat_dp(localscore: np.ndarray, period: int, tightness: float) -> Tuple[np.ndarray, np.ndarray]:
    n = len(localscore)
    backlink = np.zeros(n, dtype=np.int64)
    cumscore = np.zeros(n, dtype=localscore.dtype)
    window = np.arange(-2 * period, -int(np.round(period / 2)) + 1, dtype=np.int64)
    txwt = -tightness * (np.log(-window / period) ** 2)
    thresh = 0.01 * localscore.max()
    first = True
    nw = len(window)
    for i in range(n):
        z = int(max(0, min(-window[0], nw)))
        cand = txwt.copy()
        if z < nw:
            cand[z:] += cumscore[window[z:]]
        loc = int(np.argmax(cand))
        cumscore[i] = localscore[i] + cand[loc]
        if first and localscore[i] < thresh:
            backlink[i] = -1
        else:
            backlink[i] = window[loc]
            first = False
        window = window + 1
    return backlink, cumscore


def _localmax(x: np.ndarray) -> np.ndarray:
    m = np.zeros(len(x), dtype=bool)
    if len(x) > 2:
        m[1:-1] = (x[1:-1] > x[:-2]) & (x[1:-1] >= x[2:])
    if len(x) > 1:
        m[-1] = x[-1] > x[-2]
    return m


def beat_track(onset_envelope: np.ndarray, sr: int, hop_length: int, start_bpm: float = 120.0, tightness: float = 100.0,
               trim: bool = True, bpm: Optional[float] = None, tg: Optional[np.ndarray] = None, dp=None) -> Tuple[float, np.ndarray]:
    """Ellis dynamic-programming beat tracker: (tempo, beat frames)."""
    env = np.asarray(onset_envelope, dtype=np.float32)
    if env.size == 0 or not env.any():
        return 0.0, np.zeros(0, dtype=int)
    if bpm is None:
        if tg is None:
            tg = tempogram(env, int(np.floor(8.0 * sr / hop_length)))
        bpm = float(tempo_from_tempogram(tg, sr, hop_length, start_bpm=start_bpm)[0])
    if bpm <= 0:
        return 0.0, np.zeros(0, dtype=int)
    period = int(round(60.0 * (float(sr) / hop_length) / bpm))
    if period < 1:
        return bpm, np.zeros(0, dtype=int)
    localscore = _local_score(env, period)
    backlink, cumscore = (dp or _beat_dp)(localscore, period, tightness)
    maxes = _localmax(cumscore)
    if not maxes.any():
        return bpm, np.zeros(0, dtype=int)
    med = np.median(cumscore[maxes])
    tail = np.argwhere(cumscore * maxes * 2 > med)
    if tail.size == 0:
        return bpm, np.zeros(0, dtype=int)
    beats = [int(tail.max())]
    while backlink[beats[-1]] >= 0:
        beats.append(int(backlink[beats[-1]]))
    beats = np.array(beats[::-1], dtype=int)
    smooth = scipy.signal.convolve(env[beats], scipy.signal.windows.hann(5), "same")
    threshold = 0.5 * np.sqrt(np.mean(smooth**2)) if trim else 0.0
    valid = np.argwhere(smooth > threshold)
    if valid.size == 0:
        return bpm, np.zeros(0, dtype=int)
    return bpm, beats[int(valid.min()) : int(valid.max())]
