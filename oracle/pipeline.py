"""Oracle for the chunked separation driver and the chunk feature builder - test infrastructure.

* ``separate_track`` restates
  /root/reference/src/vocal_smart_splitter/core/enhanced_vocal_separator.py:366-373
  (chunk slicing), :423-437 (effective region accumulate) and :456-458 (uniform average).
* ``ChunkFeatures`` restates /root/reference/src/audio_cut/analysis/features_cache.py:97-101,
  122-195 (per-chunk features, effective-region mask) and :254-276, 296-297
  (first-wins dedupe, onset-frame union, MDD).  BPM / tempo / beat tracking
  (:278-294) is host logic outside the GPU path and is not restated here.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from . import features as F
from . import mdx as M
from .planner import Plan, chunk_schedule, sample_bounds


def separate_track(
    audio: np.ndarray,
    infer: Callable[[np.ndarray], tuple],
    sr: int = 44100,
    plans: Optional[Sequence[Plan]] = None,
    chunk_s=10.0,
    overlap_s=2.5,
    halo_s=0.5,
):
    """audio (N,) mono or (2,N) stereo.  ``infer(chunk) -> (vocal, instrumental)`` mono."""
    total = audio.shape[-1]
    if plans is None:
        plans = chunk_schedule(total / float(sr), chunk_s, overlap_s, halo_s)
    vacc = np.zeros(total, np.float32)
    iacc = np.zeros(total, np.float32)
    wacc = np.zeros(total, np.float32)
    for plan in plans:
        cs, ce, es, ee = sample_bounds(plan, sr, total)
        raw = audio[..., cs:ce]
        if raw.shape[-1] == 0:
            continue
        chunk = np.ascontiguousarray(raw, dtype=np.float32)
        v, ins = infer(chunk)
        ls = es - cs
        le = ls + (ee - es)
        ev = v[ls:le]
        if ev.size:
            vacc[es:ee] += ev
            wacc[es:ee] += 1.0
            iacc[es:ee] += ins[ls:le]
    wacc[wacc == 0.0] = 1.0
    vocal = (vacc / wacc).astype(np.float32)
    instrumental = (iacc / wacc).astype(np.float32) if np.any(iacc) else None
    return vocal, instrumental


def chunk_features_cpu(mix_chunk: np.ndarray, sr: int, hop: int, frame: int) -> Dict[str, np.ndarray]:
    """features_cache.py:181-195."""
    r = F.rms(mix_chunk, frame, hop)
    fl = F.spectral_flatness(mix_chunk, 2048, hop)
    on = F.onset_strength(mix_chunk, sr, hop)
    of = F.onset_detect(on, sr, hop)
    times = (np.arange(len(r)) * hop / float(sr)).astype(np.float32)
    return {"rms": r, "flat": fl, "onset_env": on, "onset_frames": of, "frame_times": times}


class ChunkFeatures:
    def __init__(self, sr: int, hop_s: float = 0.05):
        self.sr = sr
        self.hop_length = max(1, int(round(sr * hop_s)))
        self.hop_s = float(self.hop_length) / float(sr)
        self.frame_length = max(self.hop_length * 2, int(round(sr * 0.1)))
        self._rms: List[np.ndarray] = []
        self._flat: List[np.ndarray] = []
        self._onset: List[np.ndarray] = []
        self._times: List[np.ndarray] = []
        self._onset_frames: List[int] = []
        self._segments: List[np.ndarray] = []

    def add_chunk(self, plan: Plan, mix_chunk: np.ndarray, feats: Optional[Dict[str, np.ndarray]] = None) -> None:
        if mix_chunk.size == 0:
            return
        mix_chunk = np.asarray(mix_chunk, dtype=np.float32)
        if mix_chunk.ndim == 2:
            mix_chunk = np.mean(mix_chunk, axis=0)
        d = feats or chunk_features_cpu(mix_chunk, self.sr, self.hop_length, self.frame_length)
        frame_times = d["frame_times"] + plan.start_s  # float32 + python float -> float32 (NEP 50)
        eff_start = plan.start_s + plan.halo_left_s
        eff_end = plan.end_s - plan.halo_right_s
        mask = (frame_times >= eff_start) & (frame_times < eff_end)
        if not np.any(mask):
            return
        self._rms.append(d["rms"][mask])
        self._flat.append(d["flat"][mask])
        self._onset.append(d["onset_env"][mask])
        self._times.append(frame_times[mask])
        start_frame = int(round(plan.start_s / self.hop_s))
        for fr in d["onset_frames"]:
            ft = frame_times[fr] if fr < len(frame_times) else plan.start_s
            if eff_start <= ft < eff_end:
                self._onset_frames.append(start_frame + int(fr))
        es = int(round(eff_start * self.sr))
        ee = int(round(eff_end * self.sr))
        cs = int(round(plan.start_s * self.sr))
        ls = es - cs
        le = ls + (ee - es)
        if le > ls:
            self._segments.append(mix_chunk[ls:le])

    def finalize(self, w_e=0.5, w_s=0.3, w_o=0.2) -> Dict[str, np.ndarray]:
        r = np.concatenate(self._rms)
        fl = np.concatenate(self._flat)
        on = np.concatenate(self._onset)
        ft = np.concatenate(self._times)
        fidx = np.round(ft / self.hop_s).astype(int)
        uniq, first = np.unique(fidx, return_index=True)
        r = r[first].astype(np.float32)
        fl = fl[first].astype(np.float32)
        on = on[first].astype(np.float32)
        oset = set(self._onset_frames)
        onset_frames = np.array(sorted(i for i in uniq if i in oset), dtype=int)
        mdd = F.mdd_series(r, fl, on, w_e, w_s, w_o)
        return {
            "rms_series": r,
            "spectral_flatness": fl,
            "onset_envelope": on,
            "onset_frames": onset_frames,
            "frame_index": uniq,
            "mdd_series": mdd,
            "global_mdd": float(np.mean(mdd)),
            "bpm_wave": np.concatenate(self._segments) if self._segments else np.zeros(0, np.float32),
        }
