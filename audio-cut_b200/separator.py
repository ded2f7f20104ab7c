"""Separator seam - drop-in for ``EnhancedVocalSeparator``.

``B200VocalSeparator.separate_for_detection(audio, gpu_context=None) -> SeparationResult`` keeps the
contract of /root/reference/src/vocal_smart_splitter/core/enhanced_vocal_separator.py:155-205 (same
result dataclass :45-58, same ``gpu_meta`` keys) but runs the whole track in one
go: one pinned H2D of the mix, every chunk's windows batched through STFT -> U-Net -> fused iSTFT /
overlap-average on the device (``ac_separate_track``), per-chunk features in two launches, one D2H of
the stems.  When a VAD hook is installed (``vad_fn`` or a ``chunk_vad`` object with the ``SileroChunkVAD``
interface) it receives, chunk by chunk in plan order, each chunk's OWN vocal output - the whole chunk
including its halos, before trimming and overlap averaging - exactly what the reference passes at
enhanced_vocal_separator.py:412-417; the kernels emit those per-chunk stems into a side buffer
(``ac_separate_track_ex``).  There is no fallback backend: failures are recorded in ``gpu_pipeline_failures`` and
re-raised (the reference's ``strict_gpu`` behaviour).
"""
from __future__ import annotations

import os
import time
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional

import numpy as np
import torch

from . import _lib, ops
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
    return vocal_presence_markers_from_rms(rms, n, sr, hop, threshold_db, pure_music_min)


def vocal_presence_markers_from_rms(rms: np.ndarray, n: int, sr: int, hop: int, threshold_db: float = -50.0,
                                    pure_music_min: float = 0.0) -> Dict:
    """Host half of the markers: run-length segmentation of the 20 ms RMS mask and the marker cuts."""
    empty = {"vocal_presence_cut_points_sec": [], "vocal_presence_cut_points_samples": [], "vocal_presence_segments": [],
             "pure_music_segments": []}
    if sr <= 0 or n == 0:
        return empty
    duration = n / float(sr)
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


def confidence_from_energies(ve: float, ie: Optional[float], me: float) -> float:
    """``estimate_confidence`` from mean-square energies computed on the device (``ac_track_stats``)."""
    ratio = float(np.clip(ve / (me + 1e-8), 0.0, 1.0))
    if ie is not None:
        bal = ve / (ie + 1e-8)
        conf = 0.5 * ratio + 0.5 * np.clip(bal / (1.0 + bal), 0.0, 1.0)
    else:
        conf = ratio
    return float(np.clip(conf, 0.0, 1.0))


_COPY_POOL = ThreadPoolExecutor(max_workers=4, thread_name_prefix="ac-copy")


def parallel_copy(dst: np.ndarray, src: np.ndarray, min_bytes: int = 8 << 20) -> None:
    """dst[...] = src with the work split over a few threads (numpy releases the GIL while copying;
    one core moves ~6 GB/s, which would otherwise cost more than the PCIe transfer itself)."""
    if dst.shape != src.shape:
        raise ValueError(f"shape mismatch {dst.shape} vs {src.shape}")
    if dst.nbytes < min_bytes or dst.ndim == 0:
        np.copyto(dst, src, casting="same_kind")
        return
    d2 = dst.reshape(-1, dst.shape[-1]) if dst.ndim > 1 else dst.reshape(1, -1)
    s2 = src.reshape(-1, src.shape[-1]) if src.ndim > 1 else src.reshape(1, -1)
    n = d2.shape[1]
    parts = max(1, min(8, d2.nbytes // (4 << 20)))
    step = -(-n // parts)
    futs = []
    for r in range(d2.shape[0]):
        for a in range(0, n, step):
            futs.append(_COPY_POOL.submit(np.copyto, d2[r, a : a + step], s2[r, a : a + step], "same_kind"))
    for f in futs:
        f.result()


class _SmallBuf:
    def __init__(self, n: int, dev: torch.device):
        self.dev = torch.empty(n, dtype=torch.float32, device=dev)
        self.pin = torch.empty(n, dtype=torch.float32, pin_memory=True)


class _TrackViews:
    pass


class _TrackBuffers:
    """Persistent staging of one separator on one device: pinned input/output, device mix / mono / stems,
    two streams (separation; features at high priority) and the timing events.  Grown geometrically and
    reused, so a steady stream of tracks makes no allocation at all (the reference re-allocates its
    pinned staging per chunk, enhanced_vocal_separator.py:378-389)."""

    def __init__(self, dev: torch.device):
        self.dev = dev
        self.cap = 0
        self.s_sep = torch.cuda.Stream(device=dev)
        self.s_feat = torch.cuda.Stream(device=dev, priority=-1)
        self.s_copy = torch.cuda.Stream(device=dev)  # H2D pieces / D2H of finished stretches beside the window batches
        self.events = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        self.stats_dev = torch.zeros(5, dtype=torch.float64, device=dev)
        self.stats_pin = torch.zeros(5, dtype=torch.float64, pin_memory=True)
        self._small: Optional[_SmallBuf] = None

    def small(self, n: int) -> _SmallBuf:
        if self._small is None or self._small.dev.numel() < n:
            self._small = _SmallBuf(max(n, 1) * 2, self.dev)
        sb = _SmallBuf.__new__(_SmallBuf)
        sb.dev, sb.pin = self._small.dev[:n], self._small.pin[:n]
        return sb

    def views(self, n_ch: int, n: int) -> _TrackViews:
        if n > self.cap:
            cap = int(n * 1.25) + 4096
            self.pin_in = torch.empty(2 * cap, dtype=torch.float32, pin_memory=True)
            self.pin_out = torch.empty(2 * cap, dtype=torch.float32, pin_memory=True)
            self.mix = torch.empty(2 * cap, dtype=torch.float32, device=self.dev)
            self.mono = torch.empty(cap, dtype=torch.float32, device=self.dev)
            self.stems = torch.empty(3 * cap, dtype=torch.float32, device=self.dev)
            self.cap = cap
        v = _TrackViews()
        v.pin_in = self.pin_in[: n_ch * n].view(n_ch, n)
        v.pin_in_np = v.pin_in.numpy()
        v.mix = self.mix[: n_ch * n].view(n_ch, n)
        v.mono = self.mono[:n]
        st = self.stems[: 3 * n].view(3, n)
        v.vocal, v.instr, v.weight, v.stems2 = st[0], st[1], st[2], st[:2]
        v.pin_out = self.pin_out[: 2 * n].view(2, n)
        v.pin_out_np = v.pin_out.numpy()
        return v


class B200VocalSeparator:
    """``EnhancedVocalSeparator(sample_rate)`` drop-in; assign it to ``SeamlessSplitter.separator``."""

    def __init__(self, sample_rate: int = 44100, *, backend: Optional[B200Mdx23Backend] = None,
                 pipeline_config: Optional[PipelineConfig] = None,
                 vad_fn: Optional[Callable[[object, np.ndarray, int], List[Dict[str, float]]]] = None,
                 chunk_vad=None, marker_threshold_db: float = -50.0, pure_music_min_s: float = 0.0):
        self.sample_rate = sample_rate
        self._pipeline_cfg = pipeline_config or PipelineConfig()
        if backend is None:
            # enhanced_vocal_separator.py:83-137: the model directory of the checkout (env MDX23_MODELS_PATH as in the
            # reference's tests/conftest.py:25-48); no model -> RuntimeError, never silent random weights
            model_dir = os.getenv("MDX23_MODELS_PATH") or os.path.join(os.getcwd(), "MVSEP-MDX23-music-separation-model", "models")
            backend = B200Mdx23Backend(model_dir, align_hop=self._pipeline_cfg.align_hop)
        self._primary_backend: IVocalSeparatorBackend = backend
        if getattr(self._primary_backend, "_net", None) is None:
            try:
                self._primary_backend.load_model()
            except FileNotFoundError as exc:
                raise RuntimeError(f"no usable separation backend: {{'mdx23': {str(exc)!r}}}") from exc
        self._vad_fn = vad_fn
        # chunk_vad: a FACTORY ``(sample_rate) -> object with process_chunk(plan, vocal_chunk, sr) / finalize()`` (one
        # instance per track, like SileroChunkVAD at enhanced_vocal_separator.py:329-333), e.g. chunk_vad.B200ChunkVAD
        self._chunk_vad_factory = chunk_vad
        # quality_control.segment_vocal_threshold_db / quality_control.pure_music_min_duration of the reference's config
        # (vocal_separator.py:471-472); a host that carries the reference's ConfigManager passes get_config(...) here
        self._marker_threshold_db = marker_threshold_db
        self._pure_music_min_s = float(pure_music_min_s)
        self.enable_fallback = False
        self.backend_pref = "mdx23"
        # stems are returned as numpy views of page-locked memory (no pinned -> pageable copy after the D2H);
        # set False to get ordinary pageable arrays (e.g. when thousands of results are kept alive at once)
        self.pinned_outputs = True
        self.capture_device_metrics = True  # NVML / nvidia-smi sample per call (gpu_pipeline.py:208-259)
        self._bufs: Dict[torch.device, "_TrackBuffers"] = {}
        self._energies = (0.0, None, 1e-8)

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
        return SeparationResult(
            vocal_track=vocal, instrumental_track=instrumental,
            separation_confidence=confidence_from_energies(*self._energies),
            backend_used=type(backend).__name__, processing_time=time.time() - t0, quality_metrics=markers,
            feature_cache=cache, vad_segments=vad_segments, gpu_meta=ctx.to_meta(), pipeline_used=ctx.enabled,
        )

    def _separate_with_pipeline(self, audio: np.ndarray, backend: B200Mdx23Backend, ctx: PipelineContext):
        """One pinned H2D of the mix, all windows through STFT -> U-Net -> fused iSTFT/stitch, features on
        a second (high-priority) stream while the network runs, one D2H of both stems.  Every buffer is
        persistent (``_TrackBuffers``); the host blocks exactly twice: on the small feature series and on
        the final D2H event."""
        sr = self.sample_rate
        total = int(audio.shape[-1])
        dev = torch.device(ctx.device)
        bufs = self._buffers(dev)
        stream, feat_stream = bufs.s_sep, bufs.s_feat
        backend.reset_performance_metrics()
        torch.cuda.reset_peak_memory_stats(dev)
        ctx.gpu_meta.setdefault("gpu_pipeline_used", bool(ctx.enabled))
        ctx.gpu_meta.setdefault("gpu_pipeline_device", ctx.device)
        plans = ctx.plans
        bounds = [p.sample_bounds(sr, total) for p in plans]
        live = [(p, b) for p, b in zip(plans, bounds) if b[1] > b[0]]
        n_ch = 1 if audio.ndim == 1 else int(audio.shape[0])
        if n_ch not in (1, 2):
            raise ValueError(f"expected mono (N,) or stereo (2,N) audio, got shape {audio.shape}")
        v = bufs.views(n_ch, total)
        ev = bufs.events
        vad_segments: List[Dict[str, float]] = []
        lib = backend.net._lib
        out_pin = ([torch.empty(total, dtype=torch.float32, pin_memory=True) for _ in range(2)]
                   if self.pinned_outputs else None)
        finish_metrics = None
        want_chunks = self._vad_fn is not None or self._chunk_vad_factory is not None
        side_n = sum(ce - cs for _, (cs, ce, _, _) in live) if want_chunks else 0
        side_dev = torch.empty(side_n, dtype=torch.float32, device=dev) if side_n else None
        side_pin = torch.empty(side_n, dtype=torch.float32, pin_memory=True) if side_n else None
        with torch.cuda.device(dev), ctx.acquire_inflight():
            # a page-locked caller array (e.g. ``torch.empty(..., pin_memory=True).numpy()``) goes to the device as it
            # is; anything else is staged through the persistent pinned buffer first (pageable -> pinned, threaded)
            direct = (audio.dtype == np.float32 and audio.flags.c_contiguous and
                      lib.ac_host_is_pinned(audio.ctypes.data, audio.nbytes) == 1)
            if not direct:
                parallel_copy(v.pin_in_np, audio.reshape(n_ch, total))
            caller = torch.cuda.current_stream(dev)
            copy_stream = bufs.s_copy
            stream.wait_stream(caller)
            copy_stream.wait_stream(caller)
            # The copies are pipelined against the window batches (ac_separate_track_pipelined): the mix goes up in pieces,
            # each before the batch that reads it, and every stretch of the stems that no later window can touch comes
            # home while the next batch runs - only the first upload piece and the last download piece are exposed.
            host_mix = audio.ctypes.data if direct else v.pin_in_np.ctypes.data
            host_out = ((out_pin[0].data_ptr(), out_pin[1].data_ptr()) if out_pin is not None
                        else (v.pin_out[0].data_ptr(), v.pin_out[1].data_ptr()))
            with torch.cuda.stream(stream):
                ev[0].record()
            with torch.cuda.stream(copy_stream):
                ev[1].record()  # creates the CUDA event; the library re-records it once the whole mix is resident
            with torch.cuda.stream(stream):
                ops.separate_track(backend.net, v.mix, [b for _, b in live], backend.geom, align_hop=backend.align_hop,
                                   output_is_vocal=backend.get_output_type() == "vocal", dtype=backend.dtype,
                                   out=(v.vocal, v.instr, v.weight), chunk_vocal=side_dev, host_mix=host_mix, host_out=host_out,
                                   copy_stream=copy_stream, uploaded_event=ev[1])
                ev[2].record()
            feat_stream.wait_event(ev[1])
            with torch.cuda.stream(feat_stream):
                ops.check(lib.ac_downmix_mono(ops.ptr(v.mix), n_ch, total, ops.ptr(v.mono), ops.stream_ptr()), "ac_downmix_mono")
                ev[4].record()
            with torch.cuda.stream(stream):
                if self.capture_device_metrics:  # NVML sample taken while the network runs, off the critical path
                    finish_metrics = ctx.capture_device_metrics_async()
                # tail, all asynchronous: presence-marker RMS, energies, D2H of the stems and the scalars
                hop = max(1, int(0.02 * sr))
                frame = max(hop * 2, int(0.05 * sr))
                n_mark = ops.frame_count(total, frame, hop)
                mark = bufs.small(n_mark)
                ops.frame_rms(v.vocal, frame, hop, out=mark.dev)
                stream.wait_event(ev[4])
                ops.check(lib.ac_track_stats(ops.ptr(v.vocal), ops.ptr(v.instr), ops.ptr(v.mono), total, ops.ptr(bufs.stats_dev),
                                             ops.stream_ptr()), "ac_track_stats")
                if side_dev is not None:
                    side_pin.copy_(side_dev, non_blocking=True)
                mark.pin.copy_(mark.dev, non_blocking=True)
                bufs.stats_pin.copy_(bufs.stats_dev, non_blocking=True)
                ev[3].record()
            # features read the mix, not the stems: their kernels slip in between the network's on the
            # high-priority stream, and the host DSP (peak pick, beat DP) overlaps the separation
            with torch.cuda.stream(feat_stream):
                builder = B200ChunkFeatureBuilder(sr, device=str(dev))
                builder.add_track(v.mono, [p for p, _ in live])
                cache = builder.finalize(audio)
            ev[3].synchronize()
            copy_stream.synchronize()  # the last stretch of the stems
            caller.wait_stream(stream)
            caller.wait_stream(feat_stream)
            caller.wait_stream(copy_stream)
        stats = bufs.stats_pin.numpy().copy()
        if stats[4] != 0:  # tc_common.cuh:mbar_wait watchdog: the tensor-core kernels drained out early, the stems are invalid
            raise _lib.AudioCutError("a tcgen05 kernel hit its mbarrier watchdog during this track: separation output is invalid")
        any_instr = stats[3] > 0
        if out_pin is not None:
            # numpy views of page-locked tensors: freshly allocated, caller-owned (the array keeps its tensor alive;
            # torch's caching host allocator recycles the block once the caller drops the array)
            vocal = out_pin[0].numpy()
            instrumental = out_pin[1].numpy() if any_instr else None
        else:
            vocal = np.empty(total, dtype=np.float32)
            instrumental = np.empty(total, dtype=np.float32) if any_instr else None
            parallel_copy(vocal, v.pin_out_np[0])
            if instrumental is not None:
                parallel_copy(instrumental, v.pin_out_np[1])
        markers = vocal_presence_markers_from_rms(mark.pin.numpy()[:n_mark].copy(), total, sr, hop, self._marker_threshold_db,
                                                  self._pure_music_min_s)
        self._last_device = (v.mono, v.vocal, v.instr if any_instr else None)
        self._energies = (float(stats[0]) / max(total, 1), float(stats[1]) / max(total, 1) if any_instr else None,
                          float(stats[2]) / max(total, 1))
        if want_chunks:  # per-chunk VAD, in plan order, on each chunk's own output (enhanced_vocal_separator.py:412-417)
            chunk_vad = self._chunk_vad_factory(sr) if self._chunk_vad_factory is not None else None
            side_np, off = side_pin.numpy(), 0
            for p, (cs, ce, _, _) in live:
                chunk = side_np[off : off + (ce - cs)]
                off += ce - cs
                if self._vad_fn is not None:
                    vad_segments.extend(self._vad_fn(p, chunk, sr) or [])
                if chunk_vad is not None:
                    chunk_vad.process_chunk(p, chunk, sr)
            if chunk_vad is not None:
                vad_segments = vad_segments + list(chunk_vad.finalize())
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
        if finish_metrics is not None:
            finish_metrics()
        return vocal, instrumental, cache, vad_segments, markers

    def last_device_stems(self) -> Dict[str, Optional[torch.Tensor]]:
        """The last track's mono mix, vocal and instrumental stems as they lie in HBM (views of the persistent
        track buffers: valid until the next ``separate_for_detection`` call).  ``audio_cut_b200.refine.
        finalize_cut_points`` takes them as ``CutContext.mix_wave / vocal_wave`` without any copy."""
        last = getattr(self, "_last_device", None)
        if last is None:
            raise RuntimeError("no track has been separated yet")
        return {"mix": last[0], "vocal": last[1], "instrumental": last[2]}

    def _buffers(self, dev: torch.device) -> "_TrackBuffers":
        b = self._bufs.get(dev)
        if b is None:
            b = self._bufs[dev] = _TrackBuffers(dev)
        return b


EnhancedVocalSeparator = B200VocalSeparator

__all__ = ["B200VocalSeparator", "EnhancedVocalSeparator", "SeparationResult", "vocal_presence_markers", "estimate_confidence"]
