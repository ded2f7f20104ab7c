"""Oracle restatement of ``finalize_cut_points`` - test infrastructure.

Follows /root/reference/src/audio_cut/cutting/refine.py: ``nms_min_gap`` :218-250,
``align_to_zero_cross`` :72-110, ``apply_quiet_guard`` :113-158, ``_prepare_quiet_lookup`` :161-181,
``_apply_quiet_guard_fast`` :184-214, ``_filter_cut_times`` :253-266, ``finalize_cut_points`` :268-410.
north_star keeps this function on the host; it is restated here only so that the "cut points are
bit-exact in samples" claim can be checked on the GPU box, where /root/reference does not exist.
Pinned against the reference's own refine.py by tests/golden/cuts.json.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

_EPS = 1e-12


def _mono(w):
    if w is None:
        return None
    w = np.asarray(w)
    return w if w.ndim == 1 else (np.mean(w, axis=0) if w.ndim == 2 else w.reshape(-1))


def nms_min_gap(points: Sequence[Tuple[float, float]], min_gap_s: float, topk=None, max_per_window=None, window_s=10.0):
    order = sorted(range(len(points)), key=lambda i: points[i][1], reverse=True)  # stable, like sorted(..., reverse=True)
    kept: List[int] = []
    counts = {}
    span = max(window_s, min_gap_s, 1e-6)
    for i in order:
        t = points[i][0]
        if any(abs(t - points[j][0]) < min_gap_s for j in kept):
            continue
        if max_per_window is not None:
            b = int(t // span)
            if counts.get(b, 0) >= max_per_window:
                continue
            counts[b] = counts.get(b, 0) + 1
        kept.append(i)
        if topk is not None and len(kept) >= topk:
            break
    return sorted((points[i] for i in kept), key=lambda p: p[0])


def align_to_zero_cross(wave, sr: int, t: float, win_ms: float = 8.0) -> float:
    if wave is None or wave.size == 0 or sr <= 0:
        return t
    idx = int(round(t * sr))
    if idx <= 0 or idx >= wave.size:
        return t
    half = max(1, int(round(win_ms / 1000.0 * sr)))
    start, end = max(1, idx - half), min(wave.size - 1, idx + half)
    if end <= start:
        return t
    best, best_d = None, None
    for pos in range(start, end + 1):
        left, right = wave[pos - 1], wave[pos]
        if left == 0.0:
            z = pos - 1
        elif right == 0.0:
            z = pos
        elif left * right < 0.0:
            den = abs(left) + abs(right)
            z = (pos - 1) + (abs(left) / den if den > _EPS else 0.5)
        else:
            continue
        d = abs(z - idx)
        if best_d is None or d < best_d:
            best, best_d = z, d
    return t if best is None else float(best) / float(sr)


def apply_quiet_guard(wave, sr, t, *, max_shift_ms=150.0, guard_db=2.0, window_ms=10.0, floor_db=-60.0) -> float:
    if wave is None or wave.size == 0 or sr <= 0:
        return t
    idx = max(0, int(round(t * sr)))
    end = min(wave.size, idx + max(1, int(round(max_shift_ms / 1000.0 * sr))))
    if end <= idx + 1:
        return t
    seg = wave[idx:end]
    win = max(1, int(round(window_ms / 1000.0 * sr)))
    if seg.size <= win:
        rms_w = seg
    else:
        padded = np.pad(seg, (0, win - 1), mode="edge")
        rms_w = np.sqrt(np.convolve(padded * padded, np.ones(win) / float(win), mode="valid") + _EPS)
    rms_db = 20.0 * np.log10(rms_w + _EPS)
    k = int(np.argmin(rms_db))
    if (rms_db[0] - rms_db[k]) < guard_db or rms_db[k] > floor_db:
        return t
    return float(min(wave.size - 1, max(0, idx + k + win // 2))) / float(sr)


def prepare_quiet_lookup(wave, sr, window_ms, floor_db):
    if wave is None or wave.size == 0 or sr <= 0:
        return None
    win = max(1, int(round(window_ms / 1000.0 * sr)))
    sq = np.square(wave.astype(np.float64))
    rms_sq = np.convolve(sq, np.ones(win, dtype=np.float64) / float(win), mode="same")
    return 20.0 * np.log10(np.sqrt(rms_sq + _EPS) + _EPS), floor_db


def quiet_guard_fast(t, sr, lookup, *, max_shift_ms, guard_db) -> float:
    if lookup is None or sr <= 0:
        return t
    rms_db, floor_db = lookup
    n = rms_db.size
    if n == 0:
        return t
    idx = int(np.clip(int(round(t * sr)), 0, n - 1))
    end = min(n, idx + max(1, int(round(max_shift_ms / 1000.0 * sr))))
    if end <= idx:
        return t
    k = idx + int(np.argmin(rms_db[idx:end]))
    if (rms_db[idx] - rms_db[k]) < guard_db or rms_db[k] > floor_db or k == idx:
        return t
    return float(k) / float(sr)


def finalize_cut_points(mix, vocal, sr, raw_points: Sequence[Tuple[float, float]], *, use_vocal_guard_first=True,
                        min_gap_s=1.0, max_keep=None, topk_per_10s=None, nms_window_s=10.0, guard_db=2.0,
                        search_right_ms=150.0, guard_win_ms=10.0, floor_db=-60.0, enable_mix_guard=True,
                        enable_vocal_guard=True, zero_cross_win_ms=8.0, min_boundary_s=0.5):
    """raw_points: (t, score).  Returns (sample_boundaries, final_times)."""
    mix = _mono(mix)
    vocal = _mono(vocal)
    dur = len(mix) / float(sr) if sr > 0 and mix is not None else 0.0
    if mix is None or mix.size == 0 or sr <= 0:
        return [0, len(mix) if mix is not None else 0], []
    pts = list(raw_points)
    if not pts:
        return [0, len(mix)], []
    cap = topk_per_10s if (topk_per_10s is not None and topk_per_10s > 0) else None
    pruned = nms_min_gap(pts, min_gap_s, topk=max_keep, max_per_window=cap, window_s=nms_window_s)
    v_look = prepare_quiet_lookup(vocal, sr, guard_win_ms, floor_db) if enable_vocal_guard else None
    m_look = prepare_quiet_lookup(mix, sr, guard_win_ms, floor_db) if enable_mix_guard else None
    times = []
    for t, _score in pruned:
        g = t
        if use_vocal_guard_first and vocal is not None:
            g = align_to_zero_cross(vocal, sr, g, zero_cross_win_ms)
            if enable_vocal_guard:
                f = quiet_guard_fast(g, sr, v_look, max_shift_ms=search_right_ms, guard_db=guard_db)
                g = f if f != g else apply_quiet_guard(vocal, sr, g, max_shift_ms=search_right_ms, guard_db=guard_db,
                                                       window_ms=guard_win_ms, floor_db=floor_db)
        m = align_to_zero_cross(mix, sr, g, zero_cross_win_ms)
        if enable_mix_guard:
            f = quiet_guard_fast(m, sr, m_look, max_shift_ms=search_right_ms, guard_db=guard_db)
            m = f if f != m else apply_quiet_guard(mix, sr, m, max_shift_ms=search_right_ms, guard_db=guard_db,
                                                   window_ms=guard_win_ms, floor_db=floor_db)
        times.append(float(np.clip(m, 0.0, max(dur, 0.0))))
    kept: List[float] = []
    if dur > 0.0:
        boundary = min(min_boundary_s, dur / 2.0)
        for t in sorted(times):
            if t <= boundary or t >= (dur - boundary):
                continue
            if kept and (t - kept[-1]) < min_gap_s:
                continue
            kept.append(t)
    bounds = sorted(set([0] + [int(round(t * sr)) for t in kept] + [len(mix)]))
    return bounds, kept
