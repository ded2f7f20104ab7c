"""Oracle for the cut-point chain that CONSUMES the hot path's outputs - test infrastructure.

``cut_chain`` restates, for the default ``v2.2_mdd`` configuration (relative-energy-valley mode,
config/expert.yaml:38; no Silero model -> empty focus windows), what the reference does between the separator's
result and the final sample boundaries:

  PureVocalPauseDetector.detect_pure_vocal_pauses        src/vocal_smart_splitter/core/pure_vocal_pause_detector.py:131-290
      resolve_threshold                                   src/audio_cut/config/derive.py:287-327
      _estimate_vpp_multiplier                            pure_vocal_pause_detector.py:1389-1532
      _detect_energy_valleys                              :1096-1235
      _compress_pauses / _apply_total_valley_cap          :503-547 / :461-501
      _apply_mdd_enhancement (feature-cache branch)       :1237-1368
      _calculate_precise_cut_points                       :1020-1094
  candidate assembly in _process_pure_vocal_split         src/vocal_smart_splitter/core/seamless_splitter.py:435-475
      _find_no_vocal_runs                                 :1706-1790
  _finalize_and_filter_cuts_v2 -> finalize_cut_points     :1792-1877, src/audio_cut/cutting/refine.py:268-410

These are the reference's OWN host state machines (they stay on the host and are not part of the
product); they are restated here only so that a GPU test can push GPU-produced stems and series through the same chain
on a machine where /root/reference does not exist.  PINNED: ``tests/golden/cutchain.json`` is produced by
``tests/golden/make_golden.py --cutchain`` running the reference's own classes (unmodified, on the librosa shim) on
oracle stems; ``tests/test_oracle_golden.py`` requires this file to reproduce every pause, cut point, candidate and
sample boundary of that run from the same inputs.  Every configuration value the reference read during that run was
recorded by wrapping its ``get_config`` and travels in the fixture (``cfg``), so no default is guessed here.

The framewise series come from ``series`` when given (the GPU test passes the CUDA kernels' output) and from
``oracle.features`` otherwise:  ``rms_1102_441`` / ``flat_441`` (vocal, :1111-1119), ``rms_2048_441`` (vocal, :1397 and
seamless_splitter.py:1714), ``mix_rms_2048_441`` (seamless_splitter.py:1848).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import features as F


@dataclass
class Pause:
    start_time: float
    end_time: float
    duration: float
    pause_type: str
    confidence: float
    features: Dict = field(default_factory=dict)
    cut_point: float = 0.0
    quality_grade: str = "B"


class Cfg:
    """Recorded ``get_config`` answers of the reference run; a key that was never read falls back to the caller's default."""

    def __init__(self, values: Dict[str, object]):
        self.values = dict(values)

    def __call__(self, key: str, default=None):
        return self.values.get(key, default)


def _clamp(v, lo, hi):
    return max(lo, min(hi, v))


def resolve_threshold(base_ratio: float, adapt_cfg: Dict, bpm: Optional[float], global_mdd: Optional[float]):
    """derive.py:287-327 -> (peak_ratio, rms_ratio, clamp_min, clamp_max)."""
    adapt_cfg = adapt_cfg or {}
    bpm_cfg = adapt_cfg.get("bpm", {})
    cmin, cmax = float(adapt_cfg.get("clamp_min", 0.85)), float(adapt_cfg.get("clamp_max", 1.15))
    slow, fast = float(bpm_cfg.get("slow_multiplier", 1.08)), float(bpm_cfg.get("fast_multiplier", 0.92))
    peak = base_ratio
    rms = _clamp(base_ratio + 0.06, 0.05, 0.7)
    if bpm and bpm > 0:
        if bpm < 90.0:
            peak *= _clamp(slow, cmin, cmax)
        elif bpm > 140.0:
            peak *= _clamp(fast, cmin, cmax)
        peak = _clamp(peak, base_ratio * cmin, base_ratio * cmax)
    mdd_cfg = adapt_cfg.get("mdd", {})
    if global_mdd is not None:
        peak *= _clamp(float(mdd_cfg.get("base", 1.0)) + float(mdd_cfg.get("gain", 0.2)) * global_mdd, cmin, cmax)
    peak = _clamp(peak, 0.05, 0.6)
    rms = _clamp(rms, peak + 0.02, 0.72)
    return peak, rms, cmin, cmax


def _runs(mask: np.ndarray, value: bool, lo: int = 0, hi: Optional[int] = None):
    """[start, end) of the maximal runs of ``value`` inside mask[lo:hi]."""
    hi = len(mask) if hi is None else hi
    i = lo
    while i < hi:
        if bool(mask[i]) == value:
            j = i
            while j < hi and bool(mask[j]) == value:
                j += 1
            yield i, j
            i = j
        else:
            i += 1


def _close_then_open(mask: np.ndarray, close_k: int, open_k: int) -> np.ndarray:
    m = mask.astype(bool).copy()
    for a, b in list(_runs(m, False)):
        if b - a <= close_k:
            m[a:b] = True
    for a, b in list(_runs(m, True)):
        if b - a <= open_k:
            m[a:b] = False
    return m


def vpp_multiplier(rms_2048: np.ndarray, sr: int, hop: int, cfg: Cfg) -> float:
    """_estimate_vpp_multiplier (:1389-1532) without focus windows: class of the track's rest statistics."""
    db = 20.0 * np.log10(rms_2048 + 1e-12)
    delta_db = cfg("pure_vocal_detection.pause_stats_adaptation.delta_db", 3.0)
    floor_pct = float(cfg("quality_control.enforce_quiet_cut.floor_percentile", 5))
    thr = np.percentile(db, floor_pct) + float(delta_db)
    mask = db > thr
    frame_sec = hop / float(sr)
    if not np.any(mask):
        return 1.0
    close_k = max(1, int(cfg("pure_vocal_detection.pause_stats_adaptation.morph_close_ms", 150) / 1000.0 / frame_sec))
    open_k = max(1, int(cfg("pure_vocal_detection.pause_stats_adaptation.morph_open_ms", 50) / 1000.0 / frame_sec))
    mask = _close_then_open(mask, close_k, open_k)
    min_block = max(1, int(cfg("pure_vocal_detection.pause_stats_adaptation.sing_block_min_s", 2.0) / frame_sec))
    blocks = [(a, b) for a, b in _runs(mask, True) if b - a >= min_block]
    if not blocks:
        return 1.0
    interlude = int(cfg("pure_vocal_detection.pause_stats_adaptation.interlude_min_s", 4.0) / frame_sec)
    rests, total = [], 0
    for a, b in blocks:  # blocks are all-True runs, so no rest ever lies inside one; kept for fidelity with the reference
        total += b - a
        for i, j in _runs(mask, False, a, b):
            if j - i < interlude:
                rests.append((j - i) * frame_sec)
    if not rests or total == 0:
        return 1.0
    mpd, p95 = float(np.median(rests)), float(np.percentile(rests, 95))
    pr = float(len(rests) / (total * frame_sec / 60.0))
    rr = float(sum(rests) / (total * frame_sec))
    th = cfg("pure_vocal_detection.pause_stats_adaptation.classify_thresholds", {}) or {}
    s = th.get("slow", {"mpd": 0.60, "p95": 1.20, "rr": 0.35})
    f = th.get("fast", {"mpd": 0.25, "pr": 18, "rr": 0.15})
    if mpd >= s.get("mpd", 0.60) or p95 >= s.get("p95", 1.20) or rr >= s.get("rr", 0.35):
        cls = "slow"
    elif mpd <= f.get("mpd", 0.25) and pr >= f.get("pr", 18) and rr <= f.get("rr", 0.15):
        cls = "fast"
    else:
        cls = "medium"
    adapt = cfg("pure_vocal_detection.relative_threshold_adaptation", {}) or {}
    mult = adapt.get("pause_stats_multipliers") or cfg("pure_vocal_detection.relative_threshold_adaptation.pause_stats_multipliers", {}) or {}
    return float(mult.get(cls, {"slow": 1.08, "medium": 1.00, "fast": 0.92}[cls]))


def energy_valleys(rms: np.ndarray, flat: Optional[np.ndarray], sr: int, hop: int, peak_ratio: float, rms_ratio: float, cfg: Cfg) -> List[Pause]:
    """_detect_energy_valleys (:1096-1235), focus windows empty."""
    thr = min(np.max(rms) * peak_ratio, np.mean(rms) * rms_ratio)
    low = rms < thr
    times = np.arange(len(rms)) * hop / float(sr)
    w_len = cfg("pure_vocal_detection.valley_scoring.w_len", 0.6)
    w_quiet = cfg("pure_vocal_detection.valley_scoring.w_quiet", 0.4)
    w_flat = cfg("pure_vocal_detection.valley_scoring.w_flat", 0.1)
    out: List[Pause] = []
    for a, b in _runs(low, True):
        t0 = times[a]
        if b >= len(low):  # run reaches the end of the track: ends at the last frame time, fixed confidence
            t1 = times[-1]
            if t1 - t0 >= 0.2:
                out.append(Pause(t0, t1, t1 - t0, "energy_valley", 0.8, {"energy": 0.0, "threshold": thr}, (t0 + t1) / 2))
            continue
        t1 = times[b]
        dur = t1 - t0
        if dur < 0.2:
            continue
        fa, fb = max(0, int(t0 * sr / hop)), min(len(rms), int(t1 * sr / hop))
        if fa >= fb:
            continue
        energy = np.mean(rms[fa:fb])
        len_score = float(np.clip((dur - 0.20) / (1.50 - 0.20), 0.0, 1.0))
        quiet = float(np.clip(1.0 - float(energy / max(1e-12, thr)), 0.0, 1.0))
        hint = 0.5
        if flat is not None:
            sa, sb = max(0, int(t0 * sr / hop)), min(len(flat), int(t1 * sr / hop))
            if sb > sa:
                hint = float(np.clip(1.0 - float(np.mean(flat[sa:sb])), 0.0, 1.0))
        conf = max(0.1, min(0.99, w_len * len_score + w_quiet * quiet + w_flat * hint))
        out.append(Pause(t0, t1, dur, "energy_valley", conf, {"energy": energy, "threshold": thr}, (t0 + t1) / 2))
    return out


def compress_pauses(pauses: List[Pause], cfg: Cfg) -> List[Pause]:
    """_compress_pauses (:503-547)."""
    if not pauses:
        return pauses
    gap = float(cfg("pure_vocal_detection.valley_scoring.merge_close_ms", 80)) / 1000.0
    if gap > 0 and len(pauses) > 1:
        pauses = sorted(pauses, key=lambda p: p.start_time)
        merged, cur = [], pauses[0]
        for nxt in pauses[1:]:
            if nxt.start_time - cur.end_time <= gap:
                end = max(cur.end_time, nxt.end_time)
                cur = Pause(cur.start_time, end, end - cur.start_time, cur.pause_type, max(cur.confidence, nxt.confidence),
                            cur.features, 0.0, cur.quality_grade)
            else:
                merged.append(cur)
                cur = nxt
        merged.append(cur)
        pauses = merged
    max_raw = int(cfg("pure_vocal_detection.valley_scoring.max_raw_candidates", 1200))
    if len(pauses) > max_raw:
        pauses = sorted(pauses, key=lambda p: p.confidence, reverse=True)[:max_raw]
    return pauses


def total_valley_cap(pauses: List[Pause], duration_s: float, cfg: Cfg) -> List[Pause]:
    """_apply_total_valley_cap (:461-501): keep the floor(duration / segment_min_duration) quietest."""
    if not pauses:
        return pauses
    seg_min = float(cfg("quality_control.segment_min_duration", 4.0))
    if seg_min <= 0:
        seg_min = 4.0
    cap = max(1, int(math.floor(duration_s / seg_min)))
    if len(pauses) <= cap:
        return pauses

    def key(p: Pause):
        q = float(p.features.get("threshold", 0.0)) - float(p.features.get("energy", 0.0))
        return (q if np.isfinite(q) else 0.0, float(p.confidence))

    return sorted(sorted(pauses, key=key, reverse=True)[:cap], key=lambda p: p.start_time)


def mdd_enhance(pauses: List[Pause], cache, cfg: Cfg) -> List[Pause]:
    """_apply_mdd_enhancement, feature-cache branch (:1237-1368)."""
    if not pauses:
        return pauses
    rms = np.asarray(cache.rms_series, dtype=np.float32)
    flat = np.asarray(cache.spectral_flatness, dtype=np.float32)
    onset_frames = np.asarray(cache.onset_frames, dtype=np.int64)
    times = np.arange(len(rms), dtype=np.float32) * float(cache.hop_s)
    rms_max = float(cache.rms_max) if float(cache.rms_max) > 0 else 1.0
    w_e = cfg("musical_dynamic_density.energy_weight", 0.7)
    w_s = cfg("musical_dynamic_density.spectral_weight", 0.3)
    w_o = cfg("musical_dynamic_density.onset_weight", 0.2)
    t_mul = cfg("musical_dynamic_density.threshold_multiplier", 0.3)
    mx, mn = cfg("musical_dynamic_density.max_multiplier", 1.4), cfg("musical_dynamic_density.min_multiplier", 0.6)
    out = []
    for p in pauses:
        sf = int(np.argmin(np.abs(times - p.start_time))) if len(times) else 0
        ef = int(np.argmin(np.abs(times - p.end_time))) if len(times) else 0
        a, b = max(0, sf - 10), min(len(rms), ef + 10)
        if b <= a:
            out.append(p)
            continue
        idx = np.arange(a, b)
        e_score = float(np.mean(rms[idx])) / rms_max
        s_score = 1.0 - float(np.mean(flat[idx]))
        n_on = int(np.sum((onset_frames >= idx[0]) & (onset_frames <= idx[-1]))) if onset_frames.size else 0
        o_score = min(1.0, n_on / 5.0) if n_on > 0 else 0.0
        score = e_score * w_e + s_score * w_s + o_score * w_o
        mult = max(mn, min(mx, 1.0 + score * t_mul))
        out.append(Pause(p.start_time, p.end_time, p.duration, f"{p.pause_type}_mdd", p.confidence * mult,
                         {**p.features, "mdd_score": score, "confidence_multiplier": mult}, p.cut_point, p.quality_grade))
    return out


def precise_cut_points(pauses: List[Pause], vocal: np.ndarray, sr: int, cfg: Cfg) -> List[Pause]:
    """_calculate_precise_cut_points (:1020-1094): boxcar-RMS argmin, look-ahead argmin, silence-floor fallback to the
    interval midpoint."""
    win = max(1, int(float(cfg("vocal_pause_splitting.local_rms_window_ms", 25)) / 1000.0 * sr))
    guard = max(0, int(float(cfg("vocal_pause_splitting.lookahead_guard_ms", 120)) / 1000.0 * sr))
    floor_pct = float(cfg("vocal_pause_splitting.silence_floor_percentile", 5))
    allowance = float(cfg("vocal_pause_splitting.silence_floor_allowance", 1.5))

    def env(x: np.ndarray) -> np.ndarray:
        if x.size == 0:
            return np.empty(0, np.float32)
        if win <= 1:
            return np.abs(x.astype(np.float32))
        k = np.ones(win, dtype=np.float32) / float(win)
        return np.sqrt(np.maximum(np.convolve(x.astype(np.float32) ** 2, k, mode="same"), 1e-12))

    for p in pauses:
        s0 = max(0, int(round(p.start_time * sr)))
        s1 = min(len(vocal), int(round(p.end_time * sr)))
        if s1 - s0 <= 1:
            continue
        seg = vocal[s0:s1]
        cut = s0 + int(np.argmin(env(seg)))
        fallback = False
        if guard > 0:
            g1 = min(len(vocal), cut + guard)
            gseg = vocal[cut:g1]
            if gseg.size:
                cut = min(g1 - 1, cut + int(np.argmin(env(gseg))))
        floor = np.percentile(np.abs(seg), floor_pct) if seg.size else 0.0
        if floor > 0.0 and np.abs(vocal[cut]) > floor * allowance:
            cut = s0 + (s1 - s0) // 2
            fallback = True
        p.cut_point = cut / float(sr)
        p.quality_grade = "B" if fallback else "A"
    return pauses


def no_vocal_runs(rms_2048: np.ndarray, n_samples: int, sr: int, min_duration: float, cfg: Cfg) -> List[Tuple[float, float]]:
    """_find_no_vocal_runs (seamless_splitter.py:1706-1790)."""
    hop = max(1, int(0.01 * sr))
    db = 20.0 * np.log10(rms_2048 + 1e-12)
    noise_pct = float(cfg("quality_control.enforce_quiet_cut.floor_percentile", 10))
    voice_pct = float(cfg("pure_vocal_detection.pause_stats_adaptation.voice_percentile_hint", 90))
    noise_db = float(np.percentile(db, np.clip(noise_pct, 0, 50)))
    voice_db = float(np.percentile(db, np.clip(voice_pct, 50, 100)))
    delta_db = float(cfg("pure_vocal_detection.pause_stats_adaptation.delta_db", 3.0))
    thr = max(noise_db + delta_db, 0.5 * (noise_db + voice_db))
    frame_sec = hop / float(sr)
    close_k = max(1, int(int(cfg("pure_vocal_detection.pause_stats_adaptation.morph_close_ms", 150)) / 1000.0 / frame_sec))
    open_k = max(1, int(int(cfg("pure_vocal_detection.pause_stats_adaptation.morph_open_ms", 50)) / 1000.0 / frame_sec))
    inactive = ~_close_then_open(db > thr, close_k, open_k)
    times = np.arange(len(db)) * hop / float(sr)
    spans = []
    for a, b in _runs(inactive, True):
        t0 = float(times[a])
        t1 = float(times[b]) if b < len(inactive) else float(n_samples / float(sr))
        if t1 - t0 >= float(min_duration):
            spans.append((t0, t1))
    return spans


def presence_marker_times(rms_50ms: np.ndarray, n_samples: int, sr: int, cfg: Cfg) -> List[float]:
    """VocalSeparator._compute_vocal_presence_markers (src/vocal_smart_splitter/core/vocal_separator.py:460-529) ->
    'vocal_presence_cut_points_sec'; ``rms_50ms`` = librosa rms(frame max(2*hop, int(0.05 sr)), hop int(0.02 sr)) of the vocal."""
    duration = float(n_samples) / sr if n_samples > 0 else 0.0
    if n_samples == 0 or len(rms_50ms) == 0:
        return []
    thr_db = float(cfg("quality_control.segment_vocal_threshold_db", -50.0))
    pure_min = float(cfg("quality_control.pure_music_min_duration", 0.0))
    hop = max(1, int(0.02 * sr))
    mask = 20.0 * np.log10(rms_50ms + 1e-12) > thr_db
    times = np.arange(len(mask)) * hop / float(sr)
    segs, state, start = [], bool(mask[0]), 0.0
    for i in range(1, len(mask)):
        if bool(mask[i]) != state:
            segs.append((start, float(times[i]), state))
            start, state = float(times[i]), bool(mask[i])
    segs.append((start, duration, state))
    clamp = lambda v: float(min(max(v, 0.0), duration))
    cuts = set()
    first = next((g for g in segs if g[2] and g[1] > g[0]), None)
    if first is not None:
        cuts.add(clamp(first[0] - 1.0))
    for prev, nxt in zip(segs, segs[1:]):
        if not prev[2] and nxt[2] and (prev[1] - prev[0]) >= pure_min:
            c = clamp(nxt[0] - 1.0)
            if c >= prev[0]:
                cuts.add(c)
    last = next((g for g in reversed(segs) if g[2] and g[1] > g[0]), None)
    if last is not None:
        cuts.add(clamp(last[1] + 1.0))
    return sorted(c for c in cuts if 0.0 <= c <= duration)


def detect_pauses(vocal: np.ndarray, cache, cfg: Cfg, sr: int, series: Dict[str, np.ndarray]) -> List[Pause]:
    """detect_pure_vocal_pauses(vocal, enable_mdd_enhancement=True, feature_cache=cache), relative mode."""
    hop = int(sr * 0.01)
    bpm = None
    if cache.bpm_features is not None:
        bpm = float(getattr(cache.bpm_features, "main_bpm", 0.0) or 0.0)
    if bpm is not None and bpm <= 0:
        bpm = None
    mdd = float(np.clip(cache.global_mdd, 0.0, 1.0))
    peak, rms_r, cmin, cmax = resolve_threshold(cfg("pure_vocal_detection.peak_relative_threshold_ratio", 0.1),
                                               cfg("pure_vocal_detection.relative_threshold_adaptation", {}), bpm, mdd)
    if cfg("pure_vocal_detection.pause_stats_adaptation.enable", True):
        mul = float(np.clip(vpp_multiplier(series["rms_2048_441"], sr, hop, cfg), cmin, cmax))
        peak *= mul
        rms_r *= mul
    pauses = energy_valleys(series["rms_1102_441"], series.get("flat_441"), sr, hop, peak, rms_r, cfg)
    pauses = compress_pauses(pauses, cfg)
    pauses = total_valley_cap(pauses, float(len(vocal)) / float(sr), cfg)
    pauses = mdd_enhance(pauses, cache, cfg)
    if pauses:
        pauses = precise_cut_points(pauses, vocal, sr, cfg)
    return pauses


def vocal_series(vocal: np.ndarray, mix_mono: np.ndarray, sr: int) -> Dict[str, np.ndarray]:
    """The framewise series of the chain, computed by the oracle's librosa restatement."""
    hop = int(sr * 0.01)
    return {
        "rms_1102_441": F.rms(vocal, int(sr * 0.025), hop),
        "flat_441": F.spectral_flatness(vocal, 2048, hop),
        "rms_2048_441": F.rms(vocal, 2048, hop),
        "mix_rms_2048_441": F.rms(mix_mono, 2048, max(1, int(sr * 0.01))),
    }


def cut_chain(mix: np.ndarray, vocal: np.ndarray, cache, marker_times: Sequence[float], cfg_values: Dict[str, object], sr: int = 44100,
              series: Optional[Dict[str, np.ndarray]] = None, finalize: Optional[Callable] = None) -> Dict[str, object]:
    """Stems + feature cache + presence markers -> pauses, candidates and the final sample boundaries.

    ``finalize(mix, vocal, sr, points[(t, score)], **kwargs) -> (sample_boundaries, final_times)``; default = ``oracle.cuts``.
    """
    cfg = Cfg(cfg_values)
    mix_mono = mix if mix.ndim == 1 else np.mean(mix, axis=0)
    if series is None:
        series = vocal_series(vocal, mix_mono, sr)
    pauses = detect_pauses(vocal, cache, cfg, sr, series)
    cands = [(float(p.cut_point), float(p.confidence)) for p in pauses]
    min_pure = float(cfg("quality_control.pure_music_min_duration", 0.0))
    spans = no_vocal_runs(series["rms_2048_441"], len(vocal), sr, min_pure, cfg) if min_pure > 0.0 else []
    for a, b in spans:
        cands += [(float(a), 1.0), (float(b), 1.0)]
    duration = len(mix_mono) / float(sr)
    protected = set()
    for t in marker_times:
        if t <= 0.0 or t >= duration:
            continue
        cands.append((float(t), 1.0))
        protected.add(int(round(t * sr)))
    # _finalize_and_filter_cuts_v2 (seamless_splitter.py:1792-1877)
    guard_on = bool(cfg("quality_control.enforce_quiet_cut.enable", False))
    floor_db = -60.0
    if guard_on:
        override = cfg("quality_control.enforce_quiet_cut.floor_db_override", None)
        if override is not None:
            floor_db = float(override)
        else:
            fc = float(cfg("quality_control.enforce_quiet_cut.floor_percentile", 5))
            pct = fc / 100.0 if fc > 1 else fc
            rms_db = 20.0 * np.log10(series["mix_rms_2048_441"] + 1e-12)
            floor_db = float(np.percentile(rms_db, max(0.0, min(100.0, pct * 100.0))))
    topk = cfg("quality_control.nms_topk_per_10s", None)
    kw = dict(use_vocal_guard_first=True, min_gap_s=float(cfg("quality_control.min_split_gap", 1.0)),
              max_keep=int(cfg("pure_vocal_detection.valley_scoring.max_kept_after_nms", 150)),
              topk_per_10s=int(topk) if topk is not None else None, nms_window_s=float(cfg("quality_control.nms_window_s", 10.0)),
              guard_db=float(cfg("quality_control.enforce_quiet_cut.guard_db", 2.5)),
              search_right_ms=float(cfg("quality_control.enforce_quiet_cut.search_right_ms", 150)),
              guard_win_ms=float(cfg("quality_control.enforce_quiet_cut.win_ms", 80)), floor_db=floor_db,
              enable_mix_guard=guard_on, enable_vocal_guard=guard_on)
    if finalize is None:
        from . import cuts as C

        def finalize(mix_, vocal_, sr_, points, **k):
            b, t = C.finalize_cut_points(mix_, vocal_, sr_, points, **k)
            return list(b), [float(x) for x in t]

    boundaries, final_times = finalize(mix, vocal, sr, cands, **kw) if cands else ([0, len(mix_mono)], [])
    boundaries = sorted(set(int(b) for b in (boundaries or [0, len(mix_mono)])))
    total = len(mix_mono)
    final = set(boundaries)
    for s in protected:
        s = int(min(max(s, 0), total))
        if s not in (0, total):
            final.add(s)
    return {"pauses": pauses, "candidates": cands, "pure_music_spans": spans, "floor_db": floor_db, "finalize_kwargs": kw,
            "refined_boundaries": boundaries, "final_times": final_times, "sample_boundaries": sorted(final)}
