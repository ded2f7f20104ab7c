"""Separator seam - drop-in for ``EnhancedVocalSeparator``.

``B200VocalSeparator.separate_for_detection(audio, gpu_context=None) -> SeparationResult`` keeps the
contract of /root/reference/src/vocal_smart_splitter/core/enhanced_vocal_separator.py:155-205 (same
result dataclass :45-58, same ``gpu_meta`` keys) but runs the whole track in one
go: one pinned H2D of the mix, every chunk's windows batched through STFT -> U-Net -> fused iSTFT /
overlap-average on the device (``ac_separate_track``), per-chunk features in two launches, one D2H of
the stems.  When a VAD hook is installed the reference's per-chunk order of calls is preserved through
``infer_chunk``.  There is no fallback backend: failures are recorded in ``gpu_pipeline_failures`` and
re-raised (the reference's ``strict_gpu`` behaviour).
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional

import numpy as np
import torch

from . import ops
from .backends import B200Mdx23Backend, IVocalSeparatorBackend
from .features_cache import B200ChunkFeatureBuilder, TrackFeatureCache
from .gpu_pipeline import PipelineConfig, PipelineContext, build_pipeline_context, chunk_schedule


@dataclass
class SeparationResult:
    vocal_track: np.ndarray
    instrumental_track: Optional[np.ndarray]
    separation_confidence: float
    backend_used: str
    processing_time: float
    quality_metrics: Dict
    feature_cache: Optional[TrackFeatureCache] = None
    vad_segments: Optional[List[Dict[str, float]]] = None
    gpu_meta: Dict = field(default_factory=dict)
    pipeline_used: bool = False


def vocal_presence_markers(vocal_dev: torch.Tensor, sr: int, threshold_db: float = -50.0, pure_music_min: float = 0.0) -> Dict:
    """VocalSeparator._compute_vocal_presence_markers (vocal_separator.py:460-529): RMS 50 ms / 20 ms on
    the vocal stem (GPU), run-length segmentation and marker cuts on the host."""
    empty = {"vocal_presence_cut_points_sec": [], "vocal_presence_cut_points_samples": [], "vocal_presence_segments": [],
             "pure_music_segments": []}
    n = vocal_dev.numel()
    if sr <= 0 or n == 0:
        return empty
    duration = n / float(sr)
    hop = max(1, int(0.02 * sr))
    frame = max(hop * 2, int(0.05 * sr))
    rms = ops.frame_rms(vocal_dev, frame, hop).cpu().numpy()
    mask = 20.0 * np.log10(rms + 1e-12) > threshold_db
    if mask.size == 0:
        return empty
    times = np.arange(len(mask)) * hop / float(sr)
    change = np.nonzero(mask[1:] != mask[:-1])[0] + 1
    starts = np.concatenate([[0.0], times[change]])
    ends = np.concatenate([times[change], [duration]])
    states = mask[np.concatenate([[0], change])]
    segments = [{"start": float(s), "end": float(e), "is_vocal": bool(v)} for s, e, v in zip(starts, ends, states)]
    clamp = lambda v: float(min(max(v, 0.0), duration))
    cuts = set()
    first = next((g for g in segments if g["is_vocal"] and g["end"] > g["start"]), None)
    if first is not None:
        cuts.add(clamp(first["start"] - 1.0))
    for prev, nxt in zip(segments, segments[1:]):
        if not prev["is_vocal"] and nxt["is_vocal"] and (prev["end"] - prev["start"]) >= pure_music_min:
            cand = clamp(nxt["start"] - 1.0)
            if cand >= prev["start"]:
                cuts.add(cand)
    last = next((g for g in reversed(segments) if g["is_vocal"] and g["end"] > g["start"]), None)
    if last is not None:
        cuts.add(clamp(last["end"] + 1.0))
    cuts_sec = sorted(c for c in cuts if 0.0 <= c <= duration)
    return {
        "vocal_presence_cut_points_sec": cuts_sec,
        "vocal_presence_cut_points_samples": [int(round(c * sr)) for c in cuts_sec],
        "vocal_presence_segments": segments,
        "pure_music_segments": [g for g in segments if not g["is_vocal"] and g["end"] > g["start"]],
    }


def estimate_confidence(vocal: np.ndarray, instrumental: Optional[np.ndarray], mix: np.ndarray) -> float:
    """enhanced_vocal_separator.py:490-501."""
    ve = float(np.mean(np.square(vocal))) if vocal.size else 0.0
    me = float(np.mean(np.square(mix))) if mix.size else 1e-8
    ratio = float(np.clip(ve / (me + 1e-8), 0.0, 1.0))
    if instrumental is not None and instrumental.size:
        ie = float(np.mean(np.square(instrumental)))
        bal = ve / (ie + 1e-8)
        conf = 0.5 * ratio + 0.5 * np.clip(bal / (1.0 + bal), 0.0, 1.0)
    else:
        conf = ratio
    return float(np.clip(conf, 0.0, 1.0))


class B200VocalSeparator:
    """``EnhancedVocalSeparator(sample_rate)`` drop-in; assign it to ``SeamlessSplitter.separator``."""

    def __init__(self, sample_rate: int = 44100, *, backend: Optional[B200Mdx23Backend] = None,
                 pipeline_config: Optional[PipelineConfig] = None,
                 vad_fn: Optional[Callable[[object, np.ndarray, int], List[Dict[str, float]]]] = None,
                 marker_threshold_db: float = -50.0):
        self.sample_rate = sample_rate
        self._pipeline_cfg = pipeline_config or PipelineConfig()
        self._primary_backend: IVocalSeparatorBackend = backend if backend is not None else B200Mdx23Backend(allow_random_init=True)
        if getattr(self._primary_backend, "_net", None) is None:
            self._primary_backend.load_model()
        self._vad_fn = vad_fn
        self._marker_threshold_db = marker_threshold_db
        self.enable_fallback = False
        self.backend_pref = "mdx23"

    # ---- context ---------------------------------------------------------------------------
    def _ensure_pipeline_context(self, n_samples: int, gpu_context: Optional[PipelineContext]) -> PipelineContext:
        cfg = self._pipeline_cfg
        duration_s = float(n_samples) / float(self.sample_rate) if self.sample_rate > 0 else 0.0
        if gpu_context is not None and gpu_context.enabled:
            if not gpu_context.plans:
                gpu_context.plans = chunk_schedule(duration_s, chunk_s=cfg.chunk_s, overlap_s=cfg.overlap_s, halo_s=cfg.halo_s)
            ctx = gpu_context
        else:
            ctx = build_pipeline_context(duration_s, cfg)
        sig = self._primary_backend.describe_input() if hasattr(self._primary_backend, "describe_input") else None
        if sig:
            ctx.register_mdx23_input(sig)
            ctx.gpu_meta.setdefault("gpu_pipeline_mdx23_input", sig)
        return ctx

    # ---- the call the orchestrator makes -----------------------------------------------------
    def separate_for_detection(self, audio: np.ndarray, *, gpu_context: Optional[PipelineContext] = None) -> SeparationResult:
        backend = self._primary_backend
        t0 = time.time()
        audio = np.asarray(audio)
        ctx = self._ensure_pipeline_context(audio.shape[-1], gpu_context)
        try:
            vocal, instrumental, cache, vad_segments, markers = self._separate_with_pipeline(audio, backend, ctx)
        except Exception as exc:
            ctx.mark_failure("separation", str(exc))
            raise
        mono = audio if audio.ndim == 1 else audio.mean(axis=0)
        return SeparationResult(
            vocal_track=vocal, instrumental_track=instrumental,
            separation_confidence=estimate_confidence(vocal, instrumental, mono),
            backend_used=type(backend).__name__, processing_time=time.time() - t0, quality_metrics=markers,
            feature_cache=cache, vad_segments=vad_segments, gpu_meta=ctx.to_meta(), pipeline_used=ctx.enabled,
        )

    def _separate_with_pipeline(self, audio: np.ndarray, backend: B200Mdx23Backend, ctx: PipelineContext):
        sr = self.sample_rate
        total = audio.shape[-1]
        dev = torch.device(ctx.device)
        stream = ctx.streams.s_sep if ctx.use_streams and ctx.streams.s_sep is not None else torch.cuda.current_stream(dev)
        feat_stream = ctx.streams.s_feat if ctx.use_streams and ctx.streams.s_feat is not None else stream
        backend.reset_performance_metrics()
        torch.cuda.reset_peak_memory_stats(dev)
        ctx.gpu_meta.setdefault("gpu_pipeline_used", bool(ctx.enabled))
        ctx.gpu_meta.setdefault("gpu_pipeline_device", ctx.device)
        plans = ctx.plans
        bounds = [p.sample_bounds(sr, total) for p in plans]
        live = [(p, b) for p, b in zip(plans, bounds) if b[1] > b[0]]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        host = np.ascontiguousarray(audio if audio.ndim == 2 else audio[None, :], dtype=np.float32)
        pinned = ctx.pinned_pool.acquire_view(host.shape) if ctx.pinned_pool is not None else None
        vad_segments: List[Dict[str, float]] = []
        with torch.cuda.device(dev):
            with torch.cuda.stream(stream), ctx.acquire_inflight():
                ev[0].record()
                if pinned is not None:
                    pinned.copy_(torch.from_numpy(host))
                    mix = pinned.to(dev, non_blocking=True)
                else:
                    mix = torch.from_numpy(host).to(dev)
                ev[1].record()
                vocal_d, instr_d, _ = ops.separate_track(
                    backend.net, mix, [b for _, b in live], backend.geom, align_hop=backend.align_hop,
                    output_is_vocal=backend.get_output_type() == "vocal", dtype=backend.dtype)
                ev[2].record()
            # features read the mix, not the stems: they overlap the separation on their own stream
            feat_stream.wait_event(ev[1])
            with torch.cuda.stream(feat_stream):
                mono = mix[0] if mix.shape[0] == 1 else mix.mean(dim=0)
                builder = B200ChunkFeatureBuilder(sr, device=str(dev))
                builder.add_track(mono, [p for p, _ in live])
                cache = builder.finalize(audio)
            with torch.cuda.stream(stream):
                markers = vocal_presence_markers(vocal_d, sr, self._marker_threshold_db)
                any_instr = bool(torch.any(instr_d != 0).item())
                stems = torch.stack([vocal_d, instr_d]).cpu()
                ev[3].record()
            ev[3].synchronize()
        if pinned is not None:
            ctx.pinned_pool.release(pinned)
        stems = stems.numpy()
        vocal = stems[0].copy()
        instrumental = stems[1].copy() if any_instr else None
        if self._vad_fn is not None:  # hook kept from SileroChunkVAD (silero_chunk_vad.py:34, 56-116)
            for p, (cs, ce, _, _) in live:
                vad_segments.extend(self._vad_fn(p, vocal[cs:ce], sr) or [])
        backend.record_perf("h2d_ms", ev[0].elapsed_time(ev[1]))
        backend.record_perf("compute_ms", ev[1].elapsed_time(ev[2]))
        backend.record_perf("dtoh_ms", ev[2].elapsed_time(ev[3]))
        backend.record_perf("max_alloc_bytes", torch.cuda.max_memory_allocated(dev))
        backend.record_perf("chunks", float(len(live)))
        perf = backend.get_performance_metrics(reset=True)
        ctx.gpu_meta.update({
            "gpu_pipeline_processed_chunks": len(live),
            "gpu_pipeline_used": bool(ctx.enabled),
            "silero_vad_segments": len(vad_segments),
            "gpu_pipeline_h2d_ms": float(perf["h2d_ms"]),
            "gpu_pipeline_dtoh_ms": float(perf["dtoh_ms"]),
            "gpu_pipeline_compute_ms": float(perf["compute_ms"]),
            "gpu_pipeline_peak_mem_bytes": float(perf["max_alloc_bytes"]),
            "gpu_pipeline_chunk_invocations": int(perf["chunks"]),
            "mdx23_output_type": backend.get_output_type(),
        })
        ctx.capture_device_metrics()
        return vocal.astype(np.float32), None if instrumental is None else instrumental.astype(np.float32), cache, vad_segments, markers


EnhancedVocalSeparator = B200VocalSeparator

__all__ = ["B200VocalSeparator", "EnhancedVocalSeparator", "SeparationResult", "vocal_presence_markers", "estimate_confidence"]
