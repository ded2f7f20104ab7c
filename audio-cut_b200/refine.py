"""Drop-in for ``audio_cut.cutting.refine.finalize_cut_points``.

Same dataclasses, keyword arguments, return value and ordering rules as the reference
(``src/audio_cut/cutting/refine.py:16-66, 218-410``).  The score-ordered NMS (:218-250), the
minimum-gap / boundary filter (:253-266) and the bookkeeping stay on the host (a few dozen
points); the per-point signal work - zero-cross alignment and both quiet guards on the vocal
stem and on the mix - is ONE launch of ``ac_refine_cut_points`` over all pruned points.  The
waves may be the CUDA tensors the separator left in HBM (no copy) or host numpy arrays (one
upload each).  The whole-track ``QuietGuardLookup`` of the reference is never built: only the
``[idx, idx + search)`` stretch of it that a point consults is evaluated, inside the kernel.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Union

import numpy as np
import torch

from . import ops

Wave = Union[np.ndarray, torch.Tensor]


@dataclass
class CutPoint:
    t: float
    score: float
    kind: str = "pause"


@dataclass
class CutContext:
    sr: int
    mix_wave: Wave
    vocal_wave: Optional[Wave] = None


@dataclass
class CutAdjustment:
    raw_time: float
    guard_time: float
    final_time: float
    score: float
    guard_shift_ms: float
    final_shift_ms: float


@dataclass
class CutRefineResult:
    final_points: List[CutPoint]
    sample_boundaries: List[int]
    adjustments: List[CutAdjustment]
    suppressed_points: List[CutPoint] = field(default_factory=list)


def _device_mono(wave: Optional[Wave], device: torch.device) -> Optional[torch.Tensor]:
    """``_ensure_mono`` (refine.py:59-66) + residency: 1-D stays, 2-D is the float32 channel mean, else flatten."""
    if wave is None:
        return None
    if (wave.dtype if isinstance(wave, torch.Tensor) else np.asarray(wave).dtype) not in (torch.float32, np.float32):
        # the reference computes in the dtype it is given; the kernel reproduces its float32 arithmetic (the pipeline's
        # dtype: librosa.load and the separator both return float32) - refuse anything else rather than drift silently
        raise TypeError("finalize_cut_points on the GPU expects float32 waves")
    t = wave if isinstance(wave, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(wave))
    t = t.to(device=device, dtype=torch.float32, non_blocking=True)
    if t.dim() == 1:
        return t
    if t.dim() == 2:
        return t.mean(dim=0) if t.shape[0] != 2 else (t[0] + t[1]) / 2  # np.mean(axis=0) of float32 rows
    return t.reshape(-1)


def nms_min_gap(points: Iterable[CutPoint], min_gap_s: float, topk: Optional[int] = None, *,
                max_per_window: Optional[int] = None, window_s: float = 10.0) -> List[CutPoint]:
    """refine.py:218-250 - greedy by descending score (stable), minimum gap, optional per-window cap."""
    kept: List[CutPoint] = []
    counts: Dict[int, int] = {}
    span = max(window_s, min_gap_s, 1e-6)
    for point in sorted(points, key=lambda p: p.score, reverse=True):
        if any(abs(point.t - other.t) < min_gap_s for other in kept):
            continue
        bucket = None
        if max_per_window is not None:
            bucket = int(point.t // span)
            if counts.get(bucket, 0) >= max_per_window:
                continue
        kept.append(point)
        if bucket is not None:
            counts[bucket] = counts.get(bucket, 0) + 1
        if topk is not None and len(kept) >= topk:
            break
    return sorted(kept, key=lambda p: p.t)


def _filter_cut_times(times, *, duration_s: float, min_gap_s: float, min_boundary_s: float) -> List[float]:
    """refine.py:253-266."""
    out: List[float] = []
    if duration_s <= 0.0:
        return out
    boundary = min(min_boundary_s, duration_s / 2.0)
    for t in sorted(times):
        if t <= boundary or t >= (duration_s - boundary):
            continue
        if out and (t - out[-1]) < min_gap_s:
            continue
        out.append(t)
    return out


def finalize_cut_points(ctx: CutContext, raw_points: Iterable[CutPoint], *, use_vocal_guard_first: bool = True,
                        min_gap_s: float = 1.0, max_keep: Optional[int] = None, topk_per_10s: Optional[int] = None,
                        nms_window_s: float = 10.0, guard_db: float = 2.0, search_right_ms: float = 150.0,
                        guard_win_ms: float = 10.0, floor_db: float = -60.0, enable_mix_guard: bool = True,
                        enable_vocal_guard: bool = True, zero_cross_win_ms: float = 8.0,
                        min_boundary_s: float = 0.5, device: Optional[torch.device] = None) -> CutRefineResult:
    """refine.py:268-410 with the per-point loop (:318-371) on the GPU."""
    sr = ctx.sr
    if device is None:
        w = ctx.mix_wave
        device = w.device if isinstance(w, torch.Tensor) and w.is_cuda else torch.device("cuda", torch.cuda.current_device())
    mix = _device_mono(ctx.mix_wave, device)
    vocal = _device_mono(ctx.vocal_wave, device)
    n = int(mix.numel()) if mix is not None else 0
    duration_s = n / float(sr) if sr > 0 and mix is not None else 0.0
    if mix is None or n == 0 or sr <= 0:
        return CutRefineResult([], [0, n], [])
    base = list(raw_points)
    if not base:
        return CutRefineResult([], [0, n], [])
    cap = topk_per_10s if (topk_per_10s is not None and topk_per_10s > 0) else None
    pruned = nms_min_gap(base, min_gap_s=min_gap_s, topk=max_keep, max_per_window=cap, window_s=nms_window_s)
    kept_ids = {id(p) for p in pruned}
    suppressed = [CutPoint(t=float(p.t), score=float(p.score), kind=p.kind) for p in base if id(p) not in kept_ids]

    guard_t, final_t = ops.refine_cut_points(
        mix, vocal, sr, [float(p.t) for p in pruned],
        zero_cross_half=max(1, int(round(zero_cross_win_ms / 1000.0 * sr))),
        search=max(1, int(round(search_right_ms / 1000.0 * sr))),
        win=max(1, int(round(guard_win_ms / 1000.0 * sr))),
        guard_db=guard_db, floor_db=floor_db, use_vocal_guard_first=use_vocal_guard_first,
        enable_vocal_guard=enable_vocal_guard, enable_mix_guard=enable_mix_guard)

    adjustments: List[CutAdjustment] = []
    for p, g, m in zip(pruned, guard_t, final_t):
        adjustments.append(CutAdjustment(raw_time=float(p.t), guard_time=float(g), final_time=float(m), score=float(p.score),
                                         guard_shift_ms=float((g - p.t) * 1000.0), final_shift_ms=float((m - p.t) * 1000.0)))
    kept_times = _filter_cut_times([a.final_time for a in adjustments], duration_s=duration_s, min_gap_s=min_gap_s,
                                   min_boundary_s=min_boundary_s)
    kept_adj: List[CutAdjustment] = []
    for t in kept_times:
        match, best = None, None
        for adj in adjustments:
            diff = abs(adj.final_time - t)
            if best is None or diff < best:
                match, best = adj, diff
        if match is not None:
            kept_adj.append(match)
    bounds = sorted(set([0] + [int(round(t * sr)) for t in kept_times] + [n]))
    return CutRefineResult([CutPoint(t=float(t), score=1.0) for t in kept_times], bounds, kept_adj, suppressed)
