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


# ------------------------------------------------------------------------------------------------
# pYIN (librosa.pyin as called at pure_vocal_pause_detector.py:422-428) and LPC formants
# (pure_vocal_pause_detector.py:961-1018).  librosa is absent from the image: the algorithms are
# restated from the published librosa 0.10 implementation (core/pitch.py, sequence.py) - parity
# unpinned.  float64 throughout; the GPU path is compared with a tolerance.
# ------------------------------------------------------------------------------------------------
def pyin_geometry(sr=44100, fmin=65.40639132514966, fmax=2093.004522404789, frame_length=2048, win_length=None,
                  resolution=0.1):
    win_length = frame_length // 2 if win_length is None else win_length
    min_period = int(np.floor(sr / fmax))
    max_period = min(int(np.ceil(sr / fmin)), frame_length - win_length - 1)
    bins_per_semitone = int(np.ceil(1.0 / resolution))
    n_pitch_bins = int(np.floor(12 * bins_per_semitone * np.log2(fmax / fmin))) + 1
    return win_length, min_period, max_period, bins_per_semitone, n_pitch_bins


def yin_cmnd(y, frame_length=2048, win_length=1024, hop_length=441, min_period=21, max_period=675) -> np.ndarray:
    """Cumulative-mean-normalised difference, [max_period - min_period + 1, n_frames] (librosa
    _cumulative_mean_normalized_difference; the FFT autocorrelation is evaluated directly)."""
    fr = _frames(np.asarray(y, dtype=np.float64), frame_length, hop_length)  # [n_frames, frame] centred, zero pad
    n_frames = fr.shape[0]
    W = win_length
    taus = np.arange(0, max_period + 1)
    acf = np.empty((n_frames, max_period + 1))
    base = fr[:, 1:W + 1]
    for tau in taus:
        acf[:, tau] = np.sum(base * fr[:, 1 + tau:W + 1 + tau], axis=1)
    acf[np.abs(acf) < 1e-6] = 0
    cs = np.cumsum(fr ** 2, axis=1)
    energy = cs[:, W:] - cs[:, :-W]  # energy[tau] = sum_{j=tau+1}^{tau+W} y[j]^2
    energy[np.abs(energy) < 1e-6] = 0
    yin = energy[:, :1] + energy[:, :max_period + 1] - 2 * acf
    num = yin[:, min_period:max_period + 1]
    cum_mean = np.cumsum(yin[:, 1:max_period + 1], axis=1) / np.arange(1, max_period + 1)
    den = cum_mean[:, min_period - 1:max_period]
    return (num / (den + np.finfo(np.float64).tiny)).T


def parabolic_shifts(x: np.ndarray) -> np.ndarray:
    """librosa _parabolic_interpolation along axis 0 of [n, n_frames]."""
    s = np.zeros_like(x)
    a = x[2:] + x[:-2] - 2 * x[1:-1]
    b = (x[2:] - x[:-2]) / 2
    with np.errstate(divide="ignore", invalid="ignore"):
        v = np.where(np.abs(b) >= np.abs(a), 0.0, -b / a)
    s[1:-1] = v
    return s


def _beta_probs(n_thresholds=100, a=2.0, b=18.0):
    import scipy.stats

    thr = np.linspace(0, 1, n_thresholds + 1)
    return thr, np.diff(scipy.stats.beta.cdf(thr, a, b))


def pyin_observations(cmnd: np.ndarray, shifts: np.ndarray, sr=44100, fmin=65.40639132514966, min_period=21, n_pitch_bins=601,
                      bins_per_semitone=10, n_thresholds=100, boltzmann=2.0, no_trough_prob=0.01):
    """Per frame: list of (pitch bin, probability) in candidate order (increasing period; a later
    candidate overwrites an earlier one in the same bin, as numpy's fancy assignment does) and the
    voiced probability (librosa __pyin_helper)."""
    thr, beta_probs = _beta_probs(n_thresholds)
    n, n_frames = cmnd.shape
    cands = []
    voiced = np.zeros(n_frames)
    for t in range(n_frames):
        f = cmnd[:, t]
        is_trough = np.zeros(n, dtype=bool)
        is_trough[1:-1] = (f[1:-1] < f[:-2]) & (f[1:-1] <= f[2:])
        is_trough[-1] = f[-1] < f[-2]
        is_trough[0] = f[0] < f[1]
        idx = np.nonzero(is_trough)[0]
        frame_c = []
        if len(idx):
            h = f[idx]
            below = np.less.outer(h, thr[1:])                 # [troughs, thresholds]
            pos = np.cumsum(below, axis=0) - 1
            cnt = np.count_nonzero(below, axis=0)
            with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
                prior = (1 - np.exp(-boltzmann)) * np.exp(-boltzmann * pos) / (1 - np.exp(-boltzmann * cnt))
            prior = np.where(below, prior, 0.0)
            probs = prior.dot(beta_probs)
            gm = int(np.argmin(h))
            n_below_min = np.count_nonzero(~below[gm])
            probs[gm] += no_trough_prob * np.sum(beta_probs[:n_below_min])
            obs = np.zeros(n_pitch_bins + 1)
            for i, pr in zip(idx, probs):
                if pr == 0:
                    continue
                period = min_period + i + shifts[i, t]
                b = 12 * bins_per_semitone * np.log2((sr / period) / fmin)
                b = int(np.clip(np.round(b), 0, n_pitch_bins))
                frame_c.append((b, float(pr)))
                obs[b] = pr
            voiced[t] = float(np.clip(np.sum(obs[:n_pitch_bins]), 0, 1))
        cands.append(frame_c)
    return cands, voiced


def transition_local_triangle(n_states=601, width=41) -> np.ndarray:
    import scipy.signal

    win = scipy.signal.get_window("triangle", width, fftbins=False)
    T = np.zeros((n_states, n_states))
    for i in range(n_states):
        row = np.zeros(n_states)
        lpad = (n_states - width) // 2
        row[lpad:lpad + width] = win
        row = np.roll(row, n_states // 2 + i + 1)
        row[min(n_states, i + width // 2 + 1):] = 0
        row[:max(0, i - width // 2)] = 0
        T[i] = row
    return T / T.sum(axis=1, keepdims=True)


def pyin_viterbi(cands, voiced, n_pitch_bins=601, width=41, switch_prob=0.01) -> np.ndarray:
    """librosa.sequence.viterbi over 2*n_pitch_bins states with the kron(t_switch, local) transition."""
    n_steps = len(cands)
    S = 2 * n_pitch_bins
    tiny = np.finfo(np.float64).tiny
    local = transition_local_triangle(n_pitch_bins, width)
    tsw = np.array([[1 - switch_prob, switch_prob], [switch_prob, 1 - switch_prob]])
    logT = np.log(np.kron(tsw, local) + tiny)
    log_init = np.log(np.ones(S) / S + tiny)

    def log_obs(t):
        o = np.zeros(S)
        for b, pr in cands[t]:
            o[b] = pr
        o[n_pitch_bins:] = (1 - voiced[t]) / n_pitch_bins
        return np.log(o + tiny)

    value = log_obs(0) + log_init
    ptr = np.zeros((n_steps, S), dtype=np.int32)
    for t in range(1, n_steps):
        tr = value[:, None] + logT          # [from, to]
        ptr[t] = np.argmax(tr, axis=0)
        value = log_obs(t) + tr[ptr[t], np.arange(S)]
    states = np.zeros(n_steps, dtype=np.int64)
    states[-1] = int(np.argmax(value))
    for t in range(n_steps - 2, -1, -1):
        states[t] = ptr[t + 1, states[t + 1]]
    return states


def pyin(y, sr=44100, fmin=65.40639132514966, fmax=2093.004522404789, frame_length=2048, hop_length=441):
    """f0 (NaN where unvoiced), voiced_flag, voiced_prob  ==  librosa.pyin(y, fmin=C2, fmax=C7, sr, hop_length=441)."""
    W, pmin, pmax, bps, nb = pyin_geometry(sr, fmin, fmax, frame_length)
    c = yin_cmnd(y, frame_length, W, hop_length, pmin, pmax)
    sh = parabolic_shifts(c)
    cands, voiced = pyin_observations(c, sh, sr, fmin, pmin, nb, bps)
    width = int(round(35.92 * 12 * hop_length / sr)) * bps + 1
    states = pyin_viterbi(cands, voiced, nb, width)
    freqs = fmin * 2.0 ** (np.arange(nb) / (12.0 * bps))
    f0 = freqs[states % nb].astype(np.float64)
    flag = states < nb
    f0[~flag] = np.nan
    return f0, flag, voiced


def lpc_burg(y: np.ndarray, order: int) -> np.ndarray:
    """librosa.lpc (Burg's method, core/audio.py __lpc), float64; returns [1, a1..a_order]."""
    y = np.asarray(y, dtype=np.float64)
    ar = np.zeros(order + 1)
    ar[0] = 1.0
    ar_prev = ar.copy()
    fwd = y[1:].copy()
    bwd = y[:-1].copy()
    den = np.dot(fwd, fwd) + np.dot(bwd, bwd)
    eps = np.finfo(np.float64).tiny
    for i in range(order):
        rc = -2.0 * np.dot(bwd, fwd) / (den + eps)
        ar_prev, ar = ar, ar_prev
        for j in range(1, i + 2):
            ar[j] = ar_prev[j] + rc * ar_prev[i - j + 1]
        fwd_tmp = fwd
        fwd = fwd + rc * bwd
        bwd = bwd + rc * fwd_tmp
        q = 1.0 - rc ** 2
        den = q * den - bwd[-1] ** 2 - fwd[0] ** 2
        fwd = fwd[1:]
        bwd = bwd[:-1]
    return ar


def lpc_formant_frames(y, sr=44100, frame=1102, hop=441, order=12, n_freq=512):
    """Dense restatement of _extract_formants (pure_vocal_pause_detector.py:961-1018): per frame
    i in range(0, N - frame, hop): pre-emphasis 0.95, Burg LPC, |1/A| on 512 points, peaks above
    0.1*max, MAGNITUDES of the first three peaks by frequency.  Returns (mags [n_frames, 3],
    counts [n_frames]); the reference appends mags[:, j] only where count > j (ragged tracks)."""
    import scipy.signal

    y = np.asarray(y, dtype=np.float64)
    starts = list(range(0, len(y) - frame, hop))
    mags = np.zeros((len(starts), 3))
    counts = np.zeros(len(starts), dtype=np.int32)
    for r, s in enumerate(starts):
        seg = y[s:s + frame]
        pre = np.append(seg[0], seg[1:] - 0.95 * seg[:-1])
        if not np.any(pre):
            continue  # librosa.lpc raises on an all-zero frame; the reference's except: pass skips it
        a = lpc_burg(pre, order)
        if not np.all(np.isfinite(a)):
            continue
        w, h = scipy.signal.freqz([1.0], a, worN=n_freq, fs=sr)
        mag = np.abs(h)
        peaks, _ = scipy.signal.find_peaks(mag, height=float(np.max(mag)) * 0.1)
        k = min(3, len(peaks))
        mags[r, :k] = mag[peaks[:k]]
        counts[r] = k
    return mags, counts
