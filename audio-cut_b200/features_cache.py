"""Track feature cache - drop-in for ``audio_cut.analysis.features_cache``.

``TrackFeatureCache`` has the fields and helpers of the reference dataclass
(/root/reference/src/audio_cut/analysis/features_cache.py:40-91) and stays numpy-backed, because
its consumers index it on the host.  ``B200ChunkFeatureBuilder`` keeps ``ChunkFeatureBuilder``'s
interface (:94-318: ``add_chunk(plan, mix_chunk, sr, stream=)``, ``finalize(full_mix_wave)``) and
semantics - per-chunk RMS 4410/2205, flatness and mel-dB onset envelope at hop 2205 with the
top_db clip scoped to the chunk, effective-region mask, first-wins dedupe, onset-frame union, the
BPM waveform made of effective regions WITH their 1.5 s seam duplicates - with the
framewise arithmetic done by the CUDA kernels.  ``add_track`` is the batched form: every chunk of
a device-resident track in two launches.  Peak picking, tempo and beat tracking are host scans
(``host_dsp``).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import host_dsp, ops
from .gpu_pipeline import ChunkPlan

_EPS = 1e-12
MDD_WEIGHTS = (0.5, 0.3, 0.2)  # config/expert.yaml:80-83 musical_dynamic_density.{energy,spectral,onset}_weight


@dataclass
class BPMFeatures:
    """adaptive_vad_enhancer.py:16-26."""

    main_bpm: float
    bpm_category: str
    beat_strength: float
    bpm_confidence: float
    tempo_variance: float
    adaptive_factors: Dict[str, float]
    beat_positions: np.ndarray


@dataclass
class TrackFeatureCache:
    sr: int
    hop_length: int
    hop_s: float
    duration_s: float
    rms_series: np.ndarray
    spectral_flatness: np.ndarray
    onset_envelope: np.ndarray
    onset_strength: np.ndarray
    onset_frames: np.ndarray
    rms_max: float
    onset_max: float
    bpm_features: Optional[BPMFeatures]
    tempo_curve: Optional[np.ndarray]
    beat_times: np.ndarray
    global_mdd: float
    mdd_series: np.ndarray

    def frame_count(self) -> int:
        return len(self.rms_series)

    def frame_index(self, t: float) -> int:
        if self.hop_s <= 0:
            return 0
        return int(np.clip(int(round(t / self.hop_s)), 0, max(self.frame_count() - 1, 0)))

    def frame_slice(self, start_time: float, end_time: float, pad_frames: int = 0) -> slice:
        lo = max(0, self.frame_index(start_time) - pad_frames)
        hi = self.frame_index(end_time) + pad_frames + 1
        return slice(lo, min(self.frame_count(), max(lo + 1, hi)))

    def count_onsets(self, frame_slice: slice) -> int:
        if self.onset_frames.size == 0:
            return 0
        return int(np.sum((self.onset_frames >= frame_slice.start) & (self.onset_frames < frame_slice.stop)))

    def window_stats(self, start_time: float, end_time: float, pad_frames: int = 0) -> Dict[str, np.ndarray]:
        sl = self.frame_slice(start_time, end_time, pad_frames=pad_frames)
        return {"rms": self.rms_series[sl], "spectral_flatness": self.spectral_flatness[sl],
                "onset_strength": self.onset_strength[sl], "mdd": self.mdd_series[sl], "slice": sl}


def compute_mdd_series(rms, flatness, onset_strength, weights=MDD_WEIGHTS) -> np.ndarray:
    """features_cache.py:321-335."""
    w_e, w_s, w_o = weights
    r = rms / (np.max(rms) + _EPS)
    f = 1.0 - np.clip(flatness, 0.0, 1.0)
    o = onset_strength / (np.max(onset_strength) + _EPS)
    return np.clip(w_e * r + w_s * f + w_o * o, 0.0, 1.0)


def classify_bpm(bpm: float) -> str:
    """BPMAnalyzer._classify_music_by_bpm (adaptive_vad_enhancer.py:170-186)."""
    for name, lo, hi in (("slow", 50, 80), ("medium", 80, 120), ("fast", 120, 160), ("very_fast", 160, 200)):
        if lo <= bpm < hi:
            return name
    return "very_slow" if bpm < 50 else "extreme_fast"


# vocal_pause_splitting.bpm_adaptive_settings.pause_duration_multipliers.{slow,medium,fast}_song_multiplier defaults
# (adaptive_vad_enhancer.py:204-228); a host that carries the reference's ConfigManager passes its own values
PAUSE_MULTIPLIERS = (1.5, 1.0, 0.7)


def bpm_adaptive_factors(bpm: float, stability: float, variance: float, multipliers=PAUSE_MULTIPLIERS) -> Dict:
    """BPMAnalyzer._calculate_bpm_adaptive_factors / _calculate_analysis_window_size (adaptive_vad_enhancer.py:188-270)."""
    slow_m, medium_m, fast_m = multipliers
    if bpm < 70:
        f = {"threshold_modifier": -0.05, "min_pause_modifier": slow_m, "min_speech_modifier": 1.2, "sensitivity": "high"}
    elif bpm < 100:
        f = {"threshold_modifier": 0.0, "min_pause_modifier": medium_m, "min_speech_modifier": 1.0, "sensitivity": "medium"}
    elif bpm < 140:
        f = {"threshold_modifier": 0.1, "min_pause_modifier": fast_m, "min_speech_modifier": 0.8, "sensitivity": "low"}
    else:
        f = {"threshold_modifier": 0.15, "min_pause_modifier": fast_m, "min_speech_modifier": 0.6, "sensitivity": "very_low"}
    f["threshold_modifier"] += (1.0 - stability) * 0.1
    f["threshold_modifier"] += variance * 0.05
    window = 12.0 if bpm < 70 else (10.0 if bpm < 120 else 8.0)
    f.update({"bpm_value": bpm, "stability_score": stability, "variance_score": variance, "recommended_window_size": window,
              "beat_sync_important": bpm > 100})
    return f


def bpm_features_from_wave(wave_dev: torch.Tensor, sr: int) -> BPMFeatures:
    """BPMAnalyzer.extract_bpm_features (adaptive_vad_enhancer.py:48-168): onset envelope at hop 512
    with the MEDIAN aggregate on the GPU (:61-67 via beat_track(y=...), :143-148), tempogram on the GPU, the DP beat
    tracker as a C++ host scan."""
    hop = 512
    n = wave_dev.numel()
    total = 1 + n // hop
    env_d = ops.stft_features(wave_dev, [(0, n, 0)], hop, sr, total_frames=total, want=("onset_median",))["onset_median"]
    curve, bpm, _ = ops.tempogram_stats(env_d, sr, hop, start_bpm=120.0)
    env = env_d.cpu().numpy()
    if not env.any():  # librosa.beat.beat_track: no onsets -> (0.0, [])
        bpm, beats = 0.0, np.zeros(0, dtype=int)
    else:
        bpm, beats = host_dsp.beat_track(env, sr, hop, start_bpm=120.0, tightness=100.0, bpm=bpm, dp=ops.host_beat_dp)
    if len(beats) >= 3:  # _calculate_beat_stability (:99-126)
        iv = np.diff(beats)
        stability = float(np.clip(1.0 - np.std(iv) / np.mean(iv), 0.0, 1.0)) if np.mean(iv) != 0 else 0.5
    else:
        stability = 0.5
    if len(curve) > 1:  # _calculate_tempo_variance (:128-168)
        c = np.asarray(curve, dtype=np.float64)
        variance = float(np.clip(float(np.std(c)) / (float(np.mean(c)) + 1e-8), 0.0, 1.0))
    else:
        variance = 0.1
    bpm = float(bpm)
    return BPMFeatures(bpm, classify_bpm(bpm), stability, 0.8, variance, bpm_adaptive_factors(bpm, stability, variance), beats)


class B200ChunkFeatureBuilder:
    def __init__(self, sr: int, hop_s: float = 0.05, *, use_gpu: bool = True, device: Optional[str] = None) -> None:
        self.sr = sr
        self.hop_length = max(1, int(round(sr * hop_s)))
        self.hop_s = float(self.hop_length) / float(sr)
        self.frame_length = max(self.hop_length * 2, int(round(sr * 0.1)))
        self.device = torch.device(device or "cuda")
        self.use_gpu = True  # there is no other path
        self._rms: List[np.ndarray] = []
        self._flat: List[np.ndarray] = []
        self._onset: List[np.ndarray] = []
        self._times: List[np.ndarray] = []
        self._onset_frames: List[int] = []
        self._segments: List[torch.Tensor] = []  # effective-region mix pieces, device resident

    # ---- per-chunk features on the device ----------------------------------------------------
    def _chunk_features(self, mix_dev: torch.Tensor, spans: Sequence[tuple]) -> List[Dict[str, np.ndarray]]:
        """spans: (start, length) of every chunk inside ``mix_dev``; one STFT pass + one RMS launch each."""
        hop = self.hop_length
        segs, off = [], 0
        for s, l in spans:
            segs.append((s, l, off))
            off += 1 + l // hop
        feats = ops.stft_features(mix_dev, segs, hop, self.sr, total_frames=off, want=("flatness", "onset_mean"))
        # RMS frames: 1 + (l + 2*(frame//2) - frame)//hop per chunk, all chunks into one device buffer
        rms_counts = [ops.frame_count(l, self.frame_length, hop) for _, l, _ in segs]
        rms_segs, ro = [], 0
        for (s, l, _), cnt in zip(segs, rms_counts):
            rms_segs.append((s, l, ro))
            ro += cnt
        rms_all = ops.frame_rms_segments(mix_dev, rms_segs, self.frame_length, hop, total_frames=ro)  # every chunk, one launch
        packed = torch.cat([feats["flatness"], feats["onset_mean"], rms_all]).cpu().numpy()  # one D2H
        flat, onset, rms_np = packed[:off], packed[off : 2 * off], packed[2 * off :]
        out = []
        ro = 0
        for (s, l, o), cnt in zip(segs, rms_counts):
            n = 1 + l // hop
            rms = rms_np[ro : ro + cnt].copy()
            ro += cnt
            env = onset[o : o + n].copy()
            out.append({
                "rms": rms, "flat": flat[o : o + n].copy(), "onset_env": env,
                "onset_frames": host_dsp.onset_detect(env, self.sr, hop),
                "frame_times": (np.arange(len(rms)) * hop / float(self.sr)).astype(np.float32),
            })
        return out

    def _absorb(self, plan: ChunkPlan, d: Dict[str, np.ndarray], mix_chunk_dev: torch.Tensor) -> None:
        """features_cache.py:146-179: effective-region mask, onset frames to global index, BPM pieces."""
        frame_times = d["frame_times"] + plan.start_s
        eff_start, eff_end = plan.effective_start_s, plan.effective_end_s
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
        es, ee, cs = int(round(eff_start * self.sr)), int(round(eff_end * self.sr)), int(round(plan.start_s * self.sr))
        ls = es - cs
        le = ls + (ee - es)
        if le > ls:
            self._segments.append(mix_chunk_dev[ls:le])

    def add_chunk(self, plan: ChunkPlan, mix_chunk: np.ndarray, sr: int, *, stream=None) -> None:
        if mix_chunk is None or np.size(mix_chunk) == 0:
            return
        x = np.asarray(mix_chunk, dtype=np.float32)
        if x.ndim == 2:
            x = np.mean(x, axis=0)
        with torch.cuda.stream(stream if stream is not None else torch.cuda.current_stream(self.device)):
            dev = torch.from_numpy(np.ascontiguousarray(x)).to(self.device)
            d = self._chunk_features(dev, [(0, dev.numel())])[0]
            self._absorb(plan, d, dev)

    def add_track(self, mix_mono_dev: torch.Tensor, plans: Sequence[ChunkPlan]) -> None:
        """All chunks of a device-resident mono track at once (same results as add_chunk per plan)."""
        n = mix_mono_dev.numel()
        spans, kept = [], []
        for p in plans:
            cs, ce, _, _ = p.sample_bounds(self.sr, n)
            if ce > cs:
                spans.append((cs, ce - cs))
                kept.append(p)
        if not spans:
            return
        for p, (cs, ln), d in zip(kept, spans, self._chunk_features(mix_mono_dev, spans)):
            self._absorb(p, d, mix_mono_dev[cs : cs + ln])

    def finalize(self, full_mix_wave) -> TrackFeatureCache:
        if not self._rms:
            raise ValueError("no chunk produced any frame")
        rms = np.concatenate(self._rms)
        flat = np.concatenate(self._flat)
        onset = np.concatenate(self._onset)
        times = np.concatenate(self._times)
        fidx = np.round(times / self.hop_s).astype(int)
        uniq, first = np.unique(fidx, return_index=True)  # first occurrence = earlier chunk wins
        rms = rms[first].astype(np.float32, copy=False)
        flat = flat[first].astype(np.float32, copy=False)
        onset = onset[first].astype(np.float32, copy=False)
        oset = set(self._onset_frames)
        onset_frames = np.array(sorted(i for i in uniq if i in oset), dtype=int)
        bpm_wave = torch.cat(self._segments) if self._segments else None
        bpm = bpm_features_from_wave(bpm_wave, self.sr) if bpm_wave is not None and bpm_wave.numel() else None
        onset_d = torch.from_numpy(onset).to(self.device)
        tempo_curve, bpm_cache, _ = ops.tempogram_stats(onset_d, self.sr, self.hop_length)
        _, beat_frames = host_dsp.beat_track(onset, self.sr, self.hop_length, bpm=bpm_cache, dp=ops.host_beat_dp)
        beat_times = beat_frames * self.hop_length / float(self.sr)
        mdd = compute_mdd_series(rms, flat, onset)
        n_total = full_mix_wave.shape[-1] if hasattr(full_mix_wave, "shape") else len(full_mix_wave)
        return TrackFeatureCache(
            sr=self.sr, hop_length=self.hop_length, hop_s=self.hop_s, duration_s=n_total / float(self.sr),
            rms_series=rms, spectral_flatness=flat, onset_envelope=onset, onset_strength=onset.copy(),
            onset_frames=onset_frames, rms_max=float(np.max(rms) if rms.size else 0.0),
            onset_max=float(np.max(onset) if onset.size else 0.0), bpm_features=bpm, tempo_curve=tempo_curve,
            beat_times=beat_times, global_mdd=float(np.mean(mdd)), mdd_series=mdd,
        )


ChunkFeatureBuilder = B200ChunkFeatureBuilder


def build_feature_cache(mix_wave: np.ndarray, vocal_wave: Optional[np.ndarray], sr: int, *, hop_s: float = 0.05,
                        device: str = "cuda") -> TrackFeatureCache:
    """Whole-track form (features_cache.py:483-509): one segment covering the track."""
    x = np.asarray(mix_wave, dtype=np.float32)
    if x.ndim == 2:
        x = np.mean(x, axis=0)
    if x.size == 0:
        raise ValueError("mix_wave is empty, cannot build feature cache")
    b = B200ChunkFeatureBuilder(sr, hop_s, device=device)
    plan = ChunkPlan(0, 0.0, x.size / float(sr), 0.0, 0.0)
    b.add_chunk(plan, x, sr)
    return b.finalize(x)
