"""Oracle restatement of the librosa (>=0.10) functions the hot path calls - test infrastructure.

librosa is a third-party dependency (requirements.txt:5, ``librosa>=0.10.0``) that is
NOT installed in this image, so this file restates its published algorithms
 in numpy/scipy with float64 internals and float32 outputs.
PARITY UNPINNED for these functions: the reference holds no fixture for any of them.
Call sites restated:

* rms               features_cache.py:182, pure_vocal_pause_detector.py:1111-1113, 1397,
                    seamless_splitter.py:1714, 1848, vocal_separator.py:483
* spectral_flatness features_cache.py:183, pure_vocal_pause_detector.py:1117
* onset_strength    features_cache.py:184, adaptive_vad_enhancer.py:61-67, 143-148
* onset_detect      features_cache.py:186
* spectral_centroid / zero_crossing_rate / band ratio
                    pure_vocal_pause_detector.py:434-444, 937-959
"""
from __future__ import annotations

import numpy as np
import scipy.ndimage

_TINY32 = float(np.finfo(np.float32).tiny)


def frame_count(n: int, frame: int, hop: int, center: bool = True) -> int:
    padded = n + 2 * (frame // 2) if center else n
    if padded < frame:
        return 0
    return 1 + (padded - frame) // hop


def _frames(y: np.ndarray, frame: int, hop: int, pad_mode: str = "constant") -> np.ndarray:
    """(n_frames, frame) float64 view of the centered, padded signal."""
    y = np.asarray(y, dtype=np.float64)
    yp = np.pad(y, (frame // 2, frame // 2), mode=pad_mode)
    n = frame_count(len(y), frame, hop)
    idx = np.arange(n)[:, None] * hop + np.arange(frame)[None, :]
    return yp[idx]


def rms(y, frame_length=2048, hop_length=512) -> np.ndarray:
    fr = _frames(y, frame_length, hop_length)
    return np.sqrt(np.mean(fr * fr, axis=1)).astype(np.float32)


def hann(n: int) -> np.ndarray:
    return 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n) / n)  # periodic (fftbins=True)


def stft_mag2(y, n_fft=2048, hop_length=512) -> np.ndarray:
    """|STFT|^2, shape (n_frames, 1+n_fft/2); center=True, zero padding, periodic hann."""
    fr = _frames(y, n_fft, hop_length) * hann(n_fft)[None, :]
    X = np.fft.rfft(fr, axis=1)
    return X.real**2 + X.imag**2


def spectral_flatness_from_power(P: np.ndarray, amin=1e-10) -> np.ndarray:
    St = np.maximum(amin, P)
    return (np.exp(np.mean(np.log(St), axis=1)) / np.mean(St, axis=1)).astype(np.float32)


def spectral_flatness(y, n_fft=2048, hop_length=512, amin=1e-10) -> np.ndarray:
    return spectral_flatness_from_power(stft_mag2(y, n_fft, hop_length), amin)


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-12) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr=44100, n_fft=2048, n_mels=128, fmin=0.0, fmax=None) -> np.ndarray:
    """librosa.filters.mel(htk=False, norm='slaney') -> (n_mels, 1+n_fft/2) float32."""
    fmax = sr / 2.0 if fmax is None else fmax
    fftfreqs = np.linspace(0, sr / 2.0, 1 + n_fft // 2)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, 1 + n_fft // 2))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2 : n_mels + 2] - mel_f[:n_mels])
    w *= enorm[:, None]
    return w.astype(np.float32)


def power_to_db(S, amin=1e-10, top_db=80.0) -> np.ndarray:
    log_spec = 10.0 * np.log10(np.maximum(amin, S))
    if top_db is not None:
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec


def mel_db(y, sr=44100, n_fft=2048, hop_length=512, n_mels=128) -> np.ndarray:
    """power_to_db(melspectrogram(y)) with the top_db clip over the whole call; (n_frames, n_mels)."""
    P = stft_mag2(y, n_fft, hop_length)
    M = P @ mel_filterbank(sr, n_fft, n_mels).astype(np.float64).T
    return power_to_db(M)


def onset_strength(y, sr=44100, hop_length=512, n_fft=2048, aggregate=np.mean, lag=1) -> np.ndarray:
    S = mel_db(y, sr, n_fft, hop_length)  # (frames, mels)
    n = S.shape[0]
    d = np.maximum(0.0, S[lag:] - S[:-lag])
    env = aggregate(d, axis=1) if d.shape[0] else np.zeros(0)
    pad = lag + n_fft // (2 * hop_length)
    env = np.concatenate([np.zeros(pad), env])[:n]
    return env.astype(np.float32)


def peak_pick(x, pre_max, post_max, pre_avg, post_avg, delta, wait) -> np.ndarray:
    """librosa.util.peak_pick (0.10 scipy-filter formulation), written as plain loops."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    peaks = []
    last = -np.inf
    pre_max, post_max, pre_avg, post_avg, wait = (int(np.ceil(v)) for v in (pre_max, post_max, pre_avg, post_avg, wait))
    for i in range(n):
        lo, hi = max(0, i - pre_max), min(n, i + post_max)
        if x[i] != np.max(x[lo:hi]):
            continue
        lo, hi = max(0, i - pre_avg), min(n, i + post_avg)
        if x[i] < np.mean(x[lo:hi]) + delta:
            continue
        if x[i] == 0:  # detections = x * mask: zeros never survive np.nonzero
            continue
        if i > last + wait:
            peaks.append(i)
            last = i
    return np.asarray(peaks, dtype=np.int64)


def onset_detect(onset_envelope, sr=44100, hop_length=512) -> np.ndarray:
    env = np.asarray(onset_envelope)
    if not env.any() or not np.all(np.isfinite(env)):
        return np.zeros(0, dtype=np.int64)
    env = env - np.min(env)
    env = env / (np.max(env) + np.finfo(env.dtype).tiny)
    return peak_pick(
        env,
        pre_max=0.03 * sr // hop_length,
        post_max=0.00 * sr // hop_length + 1,
        pre_avg=0.10 * sr // hop_length,
        post_avg=0.10 * sr // hop_length + 1,
        delta=0.07,
        wait=0.03 * sr // hop_length,
    )


def spectral_centroid(y, sr=44100, n_fft=2048, hop_length=512) -> np.ndarray:
    S = np.sqrt(stft_mag2(y, n_fft, hop_length))
    freq = np.linspace(0, sr / 2.0, 1 + n_fft // 2)
    norm = np.maximum(S.sum(axis=1, keepdims=True), _TINY32)
    return ((S / norm) @ freq).astype(np.float32)


def low_band_ratio(y, n_fft=2048, hop_length=512) -> np.ndarray:
    """pure_vocal_pause_detector.py:937-959: sum(|S|[:n_bins//3]) / (sum(|S|) + 1e-10)."""
    S = np.sqrt(stft_mag2(y, n_fft, hop_length))
    k = S.shape[1] // 3
    return (S[:, :k].sum(axis=1) / (S.sum(axis=1) + 1e-10)).astype(np.float32)


def zero_crossing_rate(y, frame_length=2048, hop_length=512, threshold=1e-10) -> np.ndarray:
    y = np.asarray(y, dtype=np.float64)
    yp = np.pad(y, (frame_length // 2, frame_length // 2), mode="edge")
    n = frame_count(len(y), frame_length, hop_length)
    idx = np.arange(n)[:, None] * hop_length + np.arange(frame_length)[None, :]
    fr = yp[idx].copy()
    fr[np.abs(fr) <= threshold] = 0
    sb = np.signbit(fr)
    cross = sb[:, 1:] != sb[:, :-1]
    # librosa pads the crossing indicator with False at the first position
    return (cross.sum(axis=1) / float(frame_length)).astype(np.float32)


def mdd_series(rms_s, flat, onset, w_e=0.5, w_s=0.3, w_o=0.2) -> np.ndarray:
    """features_cache.py:321-335."""
    eps = 1e-12
    r = rms_s / (np.max(rms_s) + eps)
    f = 1.0 - np.clip(flat, 0.0, 1.0)
    o = onset / (np.max(onset) + eps)
    return np.clip(w_e * r + w_s * f + w_o * o, 0.0, 1.0)
