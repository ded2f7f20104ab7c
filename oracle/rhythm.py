"""Oracle for the rhythm features of TrackFeatureCache / BPMAnalyzer - test infrastructure.

Restates, in plain numpy loops (deliberately NOT the vectorised / FFT formulation the package uses in
``audio-cut_b200/host_dsp.py`` and ``csrc/features.cu``), what these reference call sites compute:

* ``librosa.feature.rhythm.tempo(onset_envelope=, sr=, hop_length=, aggregate=None)``
      /root/reference/src/audio_cut/analysis/features_cache.py:283-288   (per-frame tempo curve, hop 2205)
      /root/reference/src/vocal_smart_splitter/core/adaptive_vad_enhancer.py:151-156  (hop 512)
* ``librosa.beat.beat_track(onset_envelope=, sr=, hop_length=)``
      features_cache.py:289-294
* ``librosa.beat.beat_track(y=, sr=, hop_length=512, start_bpm=120.0, tightness=100)``
      adaptive_vad_enhancer.py:61-67   (its own onset envelope uses aggregate=np.median)
* ``BPMAnalyzer.extract_bpm_features`` and its helpers
      adaptive_vad_enhancer.py:48-168, 170-298

librosa (requirements.txt:5, ``librosa>=0.10.0``) is a third-party dependency that is absent from this
image: the algorithms are restated from the published librosa 0.10.0 / 0.10.1 sources
(``feature/rhythm.py: tempo, tempogram``, ``core/audio.py: autocorrelate``, ``util/utils.py: normalize,
localmax``, ``beat.py: beat_track, __beat_tracker, __beat_local_score, __beat_track_dp, __last_beat,
__trim_beats``).  PARITY UNPINNED against librosa itself - the reference holds no fixture for any of them
(SURVEY.md F6).  What pins this file: ``tests/test_oracle_golden.py`` checks the autocorrelation against
``np.correlate``, the tempogram normalisation / prior / argmax against closed-form click tracks (known
period -> known BPM, beats on the clicks), and the BPMAnalyzer arithmetic against
``tests/golden/rhythm.json``, which ``tests/golden/make_golden.py`` produces by running the REFERENCE's own
``BPMAnalyzer`` (adaptive_vad_enhancer.py, unmodified) on top of a librosa shim that forwards to this file -
so the classification, stability, variance and adaptive-factor arithmetic are pinned to the reference's code.

(0.10.2 rewrote the beat tracker with numba and keeps the last beat that passes the trim threshold, where
0.10.0/0.10.1 drop it - ``beats[valid.min():valid.max()]``; the earlier behaviour is the one restated, as in
round 1.  ``beat_track`` takes ``trim_inclusive=True`` to get the 0.10.2 behaviour.)
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np

from . import features as F

_TINY64 = float(np.finfo(np.float64).tiny)


def hann_periodic(n: int) -> np.ndarray:
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def autocorrelate_direct(x: np.ndarray) -> np.ndarray:
    """r[lag] = sum_n x[n] * x[n + lag], lag = 0..len(x)-1 (what librosa.autocorrelate returns, there via an FFT
    of length >= 2n-1, i.e. the LINEAR autocorrelation)."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    r = np.empty(n)
    for lag in range(n):
        r[lag] = float(np.dot(x[: n - lag], x[lag:]))
    return r


def tempogram(onset_envelope: np.ndarray, win_length: int) -> np.ndarray:
    """(win_length, n) autocorrelation tempogram: envelope padded win//2 each side with a linear ramp to 0,
    hop-1 frames, periodic hann window, linear autocorrelation, divided by max |.| per frame (frames whose
    maximum is below float tiny are left unscaled)."""
    env = np.asarray(onset_envelope, dtype=np.float64)
    n = len(env)
    half = win_length // 2
    # np.pad(mode="linear_ramp", end_values=0): ramp from 0 at the far end up to the edge value
    left = env[0] * np.arange(half) / half if half else np.zeros(0)
    right = env[-1] * (1.0 - (np.arange(half) + 1) / half) if half else np.zeros(0)
    padded = np.concatenate([left, env, right])
    window = hann_periodic(win_length)
    out = np.empty((win_length, n))
    # linear autocorrelation of every frame: correlate against a zero-extended copy, one lag at a time over all frames
    frames = np.lib.stride_tricks.sliding_window_view(padded, win_length)[:n] * window[None, :]
    for lag in range(win_length):
        out[lag] = np.einsum("ij,ij->i", frames[:, : win_length - lag], frames[:, lag:])
    mag = np.max(np.abs(out), axis=0)
    mag[mag < _TINY64] = 1.0
    return out / mag[None, :]


def tempo_frequencies(n_bins: int, sr: float, hop_length: int) -> np.ndarray:
    bpm = np.zeros(n_bins)
    bpm[0] = np.inf
    bpm[1:] = 60.0 * sr / (hop_length * np.arange(1.0, n_bins))
    return bpm


def tempo(onset_envelope: np.ndarray, sr: float, hop_length: int, start_bpm: float = 120.0, std_bpm: float = 1.0,
          ac_size: float = 8.0, max_tempo: Optional[float] = 320.0, aggregate="mean", tg: Optional[np.ndarray] = None) -> np.ndarray:
    """librosa.feature.rhythm.tempo: one value (aggregate="mean") or one per frame (aggregate=None)."""
    if tg is None:
        win_length = int(np.floor(ac_size * sr / hop_length))
        tg = tempogram(onset_envelope, win_length)
    win_length = tg.shape[0]
    if aggregate is not None:
        tg = np.mean(tg, axis=1, keepdims=True)
    bpms = tempo_frequencies(win_length, sr, hop_length)
    with np.errstate(divide="ignore"):
        logprior = -0.5 * ((np.log2(bpms) - np.log2(start_bpm)) / std_bpm) ** 2
    if max_tempo is not None:
        max_idx = int(np.argmax(bpms < max_tempo))
        logprior[:max_idx] = -np.inf
    best = np.empty(tg.shape[1], dtype=np.int64)
    for t in range(tg.shape[1]):
        score = np.log1p(1e6 * tg[:, t]) + logprior
        best[t] = int(np.argmax(score))
    return bpms[best]


def _local_score(env: np.ndarray, period: int) -> np.ndarray:
    """Gaussian-smoothed, std-normalised onset envelope (scipy.signal.convolve(..., 'same'))."""
    w = np.exp(-0.5 * (np.arange(-period, period + 1) * 32.0 / period) ** 2)
    x = env / env.std(ddof=1)
    full = np.convolve(x, w)
    start = (len(w) - 1) // 2
    return full[start : start + len(x)]


def _beat_dp(localscore: np.ndarray, period: int, tightness: float) -> Tuple[np.ndarray, np.ndarray]:
    n = len(localscore)
    backlink = np.zeros(n, dtype=np.int64)
    cumscore = np.zeros(n, dtype=localscore.dtype)
    lo, hi = -2 * period, -int(np.round(period / 2))  # previous beat offsets lo..hi inclusive
    offsets = np.arange(lo, hi + 1)
    txwt = -tightness * np.log(-offsets / period) ** 2
    thresh = 0.01 * localscore.max()
    first_beat = True
    for i in range(n):
        best_val, best_k = -np.inf, 0
        for k, off in enumerate(offsets):
            j = i + off
            v = txwt[k] + (cumscore[j] if j >= 0 else 0.0)
            if v > best_val:
                best_val, best_k = v, k
        cumscore[i] = localscore[i] + best_val
        if first_beat and localscore[i] < thresh:
            backlink[i] = -1
        else:
            backlink[i] = i + offsets[best_k]
            first_beat = False
    return backlink, cumscore


def _localmax(x: np.ndarray) -> np.ndarray:
    """librosa.util.localmax: x[i] > x[i-1] and x[i] >= x[i+1]; the first sample never, the last if > previous."""
    m = np.zeros(len(x), dtype=bool)
    for i in range(1, len(x)):
        nxt = x[i + 1] if i + 1 < len(x) else -np.inf
        m[i] = x[i] > x[i - 1] and x[i] >= nxt
    return m


def beat_track(onset_envelope: Optional[np.ndarray] = None, sr: float = 22050, hop_length: int = 512, *, y: Optional[np.ndarray] = None,
               start_bpm: float = 120.0, tightness: float = 100.0, trim: bool = True, bpm: Optional[float] = None,
               trim_inclusive: bool = False) -> Tuple[float, np.ndarray]:
    if onset_envelope is None:
        onset_envelope = F.onset_strength(y, sr, hop_length, aggregate=np.median)
    env = np.asarray(onset_envelope)
    if not env.any():
        return 0.0, np.zeros(0, dtype=int)
    if bpm is None:
        bpm = float(tempo(env, sr, hop_length, start_bpm=start_bpm)[0])
    period = int(round(60.0 * (float(sr) / hop_length) / bpm))
    localscore = _local_score(env, period)
    backlink, cumscore = _beat_dp(localscore, period, tightness)
    maxes = _localmax(cumscore)
    med = np.median(cumscore[np.argwhere(maxes)])
    last = int(np.argwhere(cumscore * maxes * 2 > med).max())
    beats = [last]
    while backlink[beats[-1]] >= 0:
        beats.append(int(backlink[beats[-1]]))
    beats = np.array(beats[::-1], dtype=int)
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(5) / 4)  # scipy.signal.hann(5), symmetric
    full = np.convolve(env[beats], w)
    smooth = full[2 : 2 + len(beats)]
    threshold = 0.5 * np.sqrt(np.mean(smooth**2)) if trim else 0.0
    valid = np.argwhere(smooth > threshold)
    return bpm, beats[int(valid.min()) : int(valid.max()) + (1 if trim_inclusive else 0)]


# ------------------------------------------------------------------------------------------------
# BPMAnalyzer (adaptive_vad_enhancer.py:27-298)
# ------------------------------------------------------------------------------------------------
@dataclass
class BPMFeatures:
    main_bpm: float
    bpm_category: str
    beat_strength: float
    bpm_confidence: float
    tempo_variance: float
    adaptive_factors: Optional[Dict] = None
    beat_positions: Optional[np.ndarray] = None


def classify_bpm(bpm: float) -> str:
    """_classify_music_by_bpm (:170-186)."""
    for name, (lo, hi) in (("slow", (50, 80)), ("medium", (80, 120)), ("fast", (120, 160)), ("very_fast", (160, 200))):
        if lo <= bpm < hi:
            return name
    return "very_slow" if bpm < 50 else "extreme_fast"


def beat_stability(beats: np.ndarray) -> float:
    """_calculate_beat_stability (:99-126)."""
    if len(beats) < 3:
        return 0.5
    iv = np.diff(beats)
    if len(iv) < 2:
        return 0.5
    m = np.mean(iv)
    if m == 0:
        return 0.5
    return float(np.clip(1.0 - np.std(iv) / m, 0.0, 1.0))


def tempo_variance(curve: np.ndarray) -> float:
    """_calculate_tempo_variance (:128-168) given the per-frame tempo curve."""
    if len(curve) > 1:
        c = np.asarray(curve, dtype=np.float64)
        return float(np.clip(float(np.std(c)) / (float(np.mean(c)) + 1e-8), 0.0, 1.0))
    return 0.1


def analysis_window_size(bpm: float) -> float:
    return 12.0 if bpm < 70 else (10.0 if bpm < 120 else 8.0)


def adaptive_factors(bpm: float, stability: float, variance: float, multipliers=(1.5, 1.0, 0.7)) -> Dict:
    """_calculate_bpm_adaptive_factors (:188-251); multipliers = (slow, medium, fast) config defaults."""
    if bpm < 70:
        f = {"threshold_modifier": -0.05, "min_pause_modifier": multipliers[0], "min_speech_modifier": 1.2, "sensitivity": "high"}
    elif bpm < 100:
        f = {"threshold_modifier": 0.0, "min_pause_modifier": multipliers[1], "min_speech_modifier": 1.0, "sensitivity": "medium"}
    elif bpm < 140:
        f = {"threshold_modifier": 0.1, "min_pause_modifier": multipliers[2], "min_speech_modifier": 0.8, "sensitivity": "low"}
    else:
        f = {"threshold_modifier": 0.15, "min_pause_modifier": multipliers[2], "min_speech_modifier": 0.6, "sensitivity": "very_low"}
    f["threshold_modifier"] += (1.0 - stability) * 0.1
    f["threshold_modifier"] += variance * 0.05
    f.update({"bpm_value": bpm, "stability_score": stability, "variance_score": variance,
              "recommended_window_size": analysis_window_size(bpm), "beat_sync_important": bpm > 100})
    return f


def extract_bpm_features(audio: np.ndarray, sr: int = 44100) -> BPMFeatures:
    """BPMAnalyzer.extract_bpm_features (:48-97) on the waveform the cache hands it (effective-region concat,
    SURVEY.md F9)."""
    hop = 512
    env = F.onset_strength(audio, sr, hop, aggregate=np.median)
    win = int(np.floor(8.0 * sr / hop))
    tg = tempogram(env, win) if env.any() else None
    if tg is None:
        bpm, beats = 0.0, np.zeros(0, dtype=int)
        curve = tempo(env, sr, hop, aggregate=None)
    else:
        bpm0 = float(tempo(env, sr, hop, start_bpm=120.0, tg=tg)[0])
        bpm, beats = beat_track(env, sr, hop, start_bpm=120.0, tightness=100.0, bpm=bpm0)
        curve = tempo(env, sr, hop, aggregate=None, tg=tg)
    stab = beat_stability(beats)
    var = tempo_variance(curve)
    return BPMFeatures(main_bpm=bpm, bpm_category=classify_bpm(bpm), beat_strength=stab, bpm_confidence=0.8, tempo_variance=var,
                       adaptive_factors=adaptive_factors(bpm, stab, var), beat_positions=beats)
