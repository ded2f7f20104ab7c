"""Per-chunk VAD adapter - drop-in for ``audio_cut.detectors.silero_chunk_vad.SileroChunkVAD``.

Keeps the adapter's contract (/root/reference/src/audio_cut/detectors/silero_chunk_vad.py:27-188):
``process_chunk(plan, vocal_chunk, sr, stream=None)`` runs the injected ``inference_fn`` on one chunk's own
vocal stem, maps its sample timestamps to track seconds, keeps what intersects the chunk's effective region
(a span that straddles the left edge keeps its early start, :103-108), ``finalize()`` merges spans closer than
``merge_gap_ms`` into ``[{start, end, duration}]`` and ``to_focus_windows`` / ``build_focus_windows`` pad and
merge them for the pause detector.  The speech model itself (Silero, third party, not in this image) stays an
injectable callable, exactly as in the reference (:34, :48-53: no model -> no segments).

What is new is the batched front end: ``process_track`` takes ALL chunks of a track at
once - the side buffer ``ac_separate_track_ex`` fills - resamples them 44.1 kHz -> 16 kHz on the GPU in one launch
(``ops.resample_chunks``: the reference resamples and pads every chunk on the host, one ``librosa.resample`` call per
chunk, core/vocal_pause_detector.py:175-296) and hands one ``[n_chunks, L]`` batch to ``batch_inference_fn``.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from .gpu_pipeline import ChunkPlan

VadFn = Callable[[np.ndarray], Sequence[Dict[str, int]]]


class B200ChunkVAD:
    def __init__(self, sample_rate: int, merge_gap_ms: float = 120.0, focus_pad_s: float = 0.2,
                 inference_fn: Optional[VadFn] = None, batch_inference_fn=None, model_rate: int = 16000):
        self.sample_rate = int(sample_rate)
        self.merge_gap_ms = float(merge_gap_ms)
        self.focus_pad_s = float(focus_pad_s)
        self.inference_fn = inference_fn
        self.batch_inference_fn = batch_inference_fn
        self.model_rate = int(model_rate)
        self._spans: List[Tuple[float, float]] = []
        self._track_end_s = 0.0
        self._final: Optional[List[Dict[str, float]]] = None

    # ---- one chunk -------------------------------------------------------------------------------
    def _absorb(self, plan: ChunkPlan, stamps: Sequence[Dict[str, int]], rate: float) -> None:
        """Timestamps in samples at ``rate`` relative to the chunk start -> clipped spans on the track timeline."""
        lo, hi = plan.effective_start_s, plan.effective_end_s
        self._track_end_s = max(self._track_end_s, float(plan.end_s))
        for ts in stamps:
            a, b = int(ts.get("start", 0)), int(ts.get("end", 0))
            if b <= a:
                continue
            t0 = plan.start_s + a / float(rate)
            t1 = plan.start_s + b / float(rate)
            if t1 <= lo or t0 >= hi:
                continue
            start = t0 if t0 < lo < t1 else max(t0, lo)  # a span crossing the left edge keeps its true onset
            end = min(t1, hi)
            if end - start > 1e-6:
                self._spans.append((start, end))
        self._spans.sort(key=lambda s: s[0])
        self._final = None

    def process_chunk(self, plan: ChunkPlan, vocal_chunk: np.ndarray, sr: int, *, stream=None) -> None:
        if np.size(vocal_chunk) == 0:
            return
        if sr != self.sample_rate:
            raise ValueError(f"B200ChunkVAD sr mismatch: expected {self.sample_rate}, got {sr}")
        fn = self.inference_fn
        if fn is None:  # no speech model installed: the reference degrades to "no segments" (:48-53)
            return
        self._absorb(plan, fn(vocal_chunk), self.sample_rate)

    # ---- all chunks of a track at once (batched front end) -----------------------------------------
    def process_track(self, plans: Sequence[ChunkPlan], chunk_vocal_dev, chunk_lens: Sequence[int]) -> None:
        """``chunk_vocal_dev``: CUDA float32 tensor with the chunks back to back (``ac_separate_track_ex``)."""
        if self.batch_inference_fn is None:
            raise RuntimeError("process_track needs batch_inference_fn")
        from . import ops

        batch, out_lens = ops.resample_chunks(chunk_vocal_dev, chunk_lens, self.sample_rate, self.model_rate)
        stamps_per_chunk = self.batch_inference_fn(batch, out_lens)  # [[{start,end} in model-rate samples], ...]
        for plan, stamps in zip(plans, stamps_per_chunk):
            self._absorb(plan, stamps, self.model_rate)

    # ---- timeline ----------------------------------------------------------------------------------
    def _merged(self) -> List[Tuple[float, float]]:
        gap = self.merge_gap_ms / 1000.0
        out: List[Tuple[float, float]] = []
        for a, b in self._spans:
            if b <= a:
                continue
            if out and a - out[-1][1] <= gap:
                out[-1] = (out[-1][0], max(out[-1][1], b))
            else:
                out.append((a, b))
        return out

    def finalize(self) -> List[Dict[str, float]]:
        if self._final is None:
            self._final = [{"start": float(a), "end": float(b), "duration": float(max(0.0, b - a))} for a, b in self._merged()]
        return list(self._final)

    def to_focus_windows(self, *, pad_s: Optional[float] = None, min_width_s: float = 0.0) -> List[Tuple[float, float]]:
        spans = self._merged() if self._final is None else [(float(e["start"]), float(e["end"])) for e in self._final]
        if not spans:
            return []
        pad = max(0.0, self.focus_pad_s if pad_s is None else float(pad_s))
        track_end = max(self._track_end_s, max(b for _, b in spans))
        wins = sorted((max(0.0, a - pad), min(track_end, b + pad)) for a, b in spans)
        out: List[Tuple[float, float]] = []
        for a, b in wins:
            if b - a <= 0.0:
                continue
            if out and a <= out[-1][1]:
                out[-1] = (out[-1][0], max(out[-1][1], b))
            else:
                out.append((a, b))
        if min_width_s > 0.0:
            out = [(a, b) for a, b in out if b - a >= min_width_s]
        return out

    def build_focus_windows(self) -> List[Tuple[float, float]]:
        return self.to_focus_windows(pad_s=self.focus_pad_s)


SileroChunkVAD = B200ChunkVAD

__all__ = ["B200ChunkVAD", "SileroChunkVAD"]
