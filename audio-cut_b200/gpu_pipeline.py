"""Chunk planning and pipeline context - host-side mirror of ``audio_cut.utils.gpu_pipeline``.

Same public names, fields and meaning as the reference module
(/root/reference/src/audio_cut/utils/gpu_pipeline.py: ``chunk_schedule`` :333-375, ``ChunkPlan``
:54-84, ``PipelineConfig`` :468-504, ``PipelineContext`` :507-577, ``PinnedBufferPool`` :378-421,
``InflightLimiter`` :428-465, ``build_pipeline_context`` :580-642) so that objects built here can be
handed to the reference's orchestrator (``SeamlessSplitter._build_gpu_pipeline_context``) and the
``gpu_pipeline_*`` manifest keys come out the same.  Differences, on purpose:
there is no onnxruntime section (``ort`` keys are accepted and ignored) and no CPU context - the
device is always a CUDA device and ``strict_gpu`` is effectively always on.
"""
from __future__ import annotations

import threading
from contextlib import contextmanager
from dataclasses import dataclass, field
from typing import Dict, Iterator, List, Optional, Sequence

import torch


@dataclass
class Streams:
    s_sep: Optional["torch.cuda.Stream"] = None
    s_vad: Optional["torch.cuda.Stream"] = None
    s_feat: Optional["torch.cuda.Stream"] = None

    def as_tuple(self):
        return (self.s_sep, self.s_vad, self.s_feat)


@dataclass
class ChunkPlan:
    index: int
    start_s: float
    end_s: float
    halo_left_s: float
    halo_right_s: float

    @property
    def duration_s(self) -> float:
        return max(0.0, self.end_s - self.start_s)

    @property
    def effective_start_s(self) -> float:
        return self.start_s + self.halo_left_s

    @property
    def effective_end_s(self) -> float:
        return self.end_s - self.halo_right_s

    def as_slice(self, sample_rate: int) -> slice:
        lo = max(0, int(round(self.start_s * sample_rate)))
        return slice(lo, max(lo, int(round(self.end_s * sample_rate))))

    def sample_bounds(self, sample_rate: int, total_samples: int):
        """(chunk_start, chunk_end, eff_start, eff_end) exactly as
        enhanced_vocal_separator.py:367-368, 423-425 round them."""
        cs = max(0, int(round(self.start_s * sample_rate)))
        ce = min(total_samples, int(round(self.end_s * sample_rate)))
        es = cs + int(round(self.halo_left_s * sample_rate))
        ee = ce - int(round(self.halo_right_s * sample_rate))
        ee = max(es, min(total_samples, ee))
        return cs, ce, es, ee


def chunk_schedule(total_s: float, *, chunk_s: float = 10.0, overlap_s: float = 2.5, halo_s: float = 0.5) -> List[ChunkPlan]:
    """Chunks of ``chunk_s`` every ``chunk_s - overlap_s`` seconds; inner edges carry a halo.

    Float stepping (``start += stride``) and the 1e-6 guards follow the reference so that plans -
    and therefore every sample bound derived from them - are bit-identical.
    """
    total_s = max(0.0, float(total_s))
    chunk_s = max(0.1, float(chunk_s))
    overlap_s = max(0.0, min(float(overlap_s), 0.9 * chunk_s))
    halo_s = max(0.0, min(float(halo_s), 0.5 * chunk_s))
    if total_s <= chunk_s:
        return [ChunkPlan(0, 0.0, total_s, 0.0, 0.0)]
    stride = chunk_s - overlap_s
    if stride <= 0:
        stride = chunk_s
    plans: List[ChunkPlan] = []
    start = 0.0
    while start < total_s - 1e-6:
        end = min(total_s, start + chunk_s)
        more = end < total_s - 1e-6
        plans.append(ChunkPlan(len(plans), start, end, halo_s if plans else 0.0, halo_s if more else 0.0))
        if not more:
            break
        start += stride
    return plans


def select_device(preferred: Optional[str] = None) -> str:
    if not torch.cuda.is_available():
        raise RuntimeError("audio_cut_b200 needs a CUDA device (no CPU fallback)")
    name = (preferred or "cuda").strip().lower() or "cuda"
    if name in ("cuda", "gpu"):
        return f"cuda:{torch.cuda.current_device()}"
    if ":" in name:
        name = name.split(":", 1)[1]
    try:
        idx = int(name)
    except ValueError:
        idx = 0
    if not 0 <= idx < torch.cuda.device_count():
        idx = 0
    return f"cuda:{idx}"


def create_streams(device: str, enable: bool = True) -> Streams:
    if not enable:
        return Streams()
    dev = torch.device(device)
    return Streams(*(torch.cuda.Stream(device=dev) for _ in range(3)))


def record_event(stream, *, enable_timing: bool = False):
    if stream is None:
        return None
    ev = torch.cuda.Event(enable_timing=enable_timing)
    ev.record(stream)
    return ev


def wait_event(stream, event) -> None:
    if stream is not None and event is not None:
        stream.wait_event(event)


@dataclass
class PinnedBufferPool:
    """Page-locked host staging buffers, reused across tracks (capacity buffers are kept)."""

    dtype: "torch.dtype" = torch.float32
    capacity: int = 2
    _free: List["torch.Tensor"] = field(default_factory=list)

    def acquire(self, num_elements: int):
        if num_elements <= 0:
            return None
        for i, buf in enumerate(self._free):
            if buf.numel() >= num_elements:
                return self._free.pop(i)[:num_elements]
        return torch.empty(int(num_elements), dtype=self.dtype, pin_memory=True)

    def acquire_view(self, shape: Sequence[int]):
        n = 1
        for d in shape:
            n *= int(d)
        t = self.acquire(n)
        return None if t is None else t.view(*shape)

    def release(self, tensor) -> None:
        if tensor is None:
            return
        base = tensor.reshape(-1)
        if len(self._free) < self.capacity:
            self._free.append(base)

    def clear(self) -> None:
        self._free.clear()


@dataclass
class InflightLimiter:
    """Counting gate: at most ``limit`` holders at a time (0 = unlimited)."""

    limit: int
    _cv: threading.Condition = field(default_factory=threading.Condition, init=False)
    _n: int = field(default=0, init=False)

    def __post_init__(self):
        self.limit = max(0, int(self.limit))

    @contextmanager
    def acquire(self, timeout: Optional[float] = None) -> Iterator[None]:
        if self.limit == 0:
            yield
            return
        with self._cv:
            if not self._cv.wait_for(lambda: self._n < self.limit, timeout=timeout):
                raise RuntimeError("inflight limit exceeded")
            self._n += 1
        try:
            yield
        finally:
            with self._cv:
                self._n = max(0, self._n - 1)
                self._cv.notify()


@dataclass
class PipelineConfig:
    enable: bool = True
    prefer_device: str = "cuda"
    chunk_s: float = 10.0
    overlap_s: float = 2.5
    halo_s: float = 0.5
    align_hop: int = 4096
    use_cuda_streams: bool = True
    prefetch_pinned_buffers: int = 2
    inflight_chunks_limit: int = 2
    strict_gpu: bool = True

    @classmethod
    def from_mapping(cls, mapping: Optional[dict]) -> "PipelineConfig":
        m = mapping or {}
        return cls(
            enable=bool(m.get("enable", True)),
            prefer_device=str(m.get("prefer_device", "cuda")),
            chunk_s=float(m.get("chunk_seconds", m.get("chunk_s", 10.0))),
            overlap_s=float(m.get("overlap_seconds", m.get("overlap_s", 2.5))),
            halo_s=float(m.get("halo_seconds", m.get("halo_s", 0.5))),
            align_hop=int(m.get("align_hop", m.get("align_hop_samples", 4096))),
            use_cuda_streams=bool(m.get("use_cuda_streams", True)),
            prefetch_pinned_buffers=int(m.get("prefetch_pinned_buffers", 2)),
            inflight_chunks_limit=int(m.get("inflight_chunks_limit", 2)),
            strict_gpu=bool(m.get("strict_mode", m.get("strict_gpu", True))),
        )


@dataclass
class PipelineContext:
    device: str
    streams: Streams
    plans: List[ChunkPlan]
    pinned_pool: Optional[PinnedBufferPool]
    limiter: Optional[InflightLimiter]
    config: PipelineConfig = field(repr=False)
    use_streams: bool = False
    strict_gpu: bool = True
    mdx23_input: Optional[Dict[str, List[int]]] = None
    gpu_meta: Dict[str, object] = field(default_factory=dict)
    failures: List[Dict[str, str]] = field(default_factory=list)
    device_index: Optional[int] = None
    device_name: Optional[str] = None

    @property
    def enabled(self) -> bool:
        return bool(self.config.enable and str(self.device).startswith("cuda") and self.use_streams and self.streams.s_sep)

    @contextmanager
    def acquire_inflight(self, timeout: Optional[float] = None) -> Iterator[None]:
        if self.limiter is None:
            yield
        else:
            with self.limiter.acquire(timeout=timeout):
                yield

    def register_mdx23_input(self, info) -> None:
        self.mdx23_input = info

    def mark_failure(self, stage: str, reason: str) -> None:
        self.failures.append({"stage": stage, "reason": reason})

    def to_meta(self) -> Dict[str, object]:
        meta = dict(self.gpu_meta)
        defaults = {
            "gpu_pipeline_enabled": bool(self.config.enable),
            "gpu_pipeline_used": bool(self.enabled),
            "gpu_pipeline_device": self.device,
            "gpu_pipeline_chunks": len(self.plans),
            "gpu_pipeline_streams": bool(self.use_streams),
            "gpu_pipeline_inflight_limit": int(self.limiter.limit) if self.limiter else 0,
            "gpu_pipeline_prefetch": int(self.pinned_pool.capacity) if self.pinned_pool else 0,
            "gpu_pipeline_align_hop": int(self.config.align_hop),
            "gpu_pipeline_config": {
                "chunk_seconds": float(self.config.chunk_s),
                "overlap_seconds": float(self.config.overlap_s),
                "halo_seconds": float(self.config.halo_s),
            },
        }
        if self.device_index is not None:
            defaults["gpu_pipeline_device_index"] = int(self.device_index)
        if self.device_name:
            defaults["gpu_pipeline_device_name"] = self.device_name
        if self.mdx23_input:
            defaults["gpu_pipeline_mdx23_input"] = self.mdx23_input
        if self.failures:
            defaults["gpu_pipeline_failures"] = list(self.failures)
        for k, v in defaults.items():
            meta.setdefault(k, v)
        return meta

    def capture_device_metrics_async(self):
        """Start the NVML sample on a helper thread WHILE the kernels run (the sample then shows the device
        under load, and the 10+ ms NVML round trip leaves the caller's critical path); returns a callable
        that joins the thread and merges the gpu_pipeline_nvml_* keys into ``gpu_meta``."""
        import threading

        th = threading.Thread(target=self.capture_device_metrics, daemon=True)
        th.start()

        def finish(timeout: float = 2.0) -> None:
            th.join(timeout)

        return finish

    def capture_device_metrics(self) -> None:
        """gpu_pipeline_nvml_* keys (gpu_pipeline.py:208-259) through pynvml when it is importable."""
        try:
            import pynvml  # nvidia-ml-py

            global _NVML_READY
            if not _NVML_READY:
                pynvml.nvmlInit()
                _NVML_READY = True
            h = pynvml.nvmlDeviceGetHandleByIndex(int(self.device_index or 0))
            util = pynvml.nvmlDeviceGetUtilizationRates(h)
            mem = pynvml.nvmlDeviceGetMemoryInfo(h)
            self.gpu_meta.update(
                {
                    "gpu_pipeline_nvml_gpu_util_percent": float(util.gpu),
                    "gpu_pipeline_nvml_mem_util_percent": float(util.memory),
                    "gpu_pipeline_nvml_mem_used_bytes": float(mem.used),
                    "gpu_pipeline_nvml_mem_total_bytes": float(mem.total),
                }
            )
        except Exception:
            pass


_NVML_READY = False


def build_pipeline_context(duration_s: float, cfg: PipelineConfig) -> PipelineContext:
    device = select_device(cfg.prefer_device)
    dev = torch.device(device)
    torch.cuda.set_device(dev)
    with torch.cuda.device(dev):
        streams = create_streams(device, cfg.use_cuda_streams)
    plans = chunk_schedule(duration_s, chunk_s=cfg.chunk_s, overlap_s=cfg.overlap_s, halo_s=cfg.halo_s)
    pool = PinnedBufferPool(dtype=torch.float32, capacity=max(1, cfg.prefetch_pinned_buffers)) if cfg.prefetch_pinned_buffers > 0 else None
    limiter = InflightLimiter(limit=cfg.inflight_chunks_limit) if cfg.inflight_chunks_limit > 0 else None
    ctx = PipelineContext(
        device=device, streams=streams, plans=plans, pinned_pool=pool, limiter=limiter, config=cfg,
        use_streams=bool(cfg.use_cuda_streams), strict_gpu=True, device_index=dev.index,
        device_name=torch.cuda.get_device_name(dev),
    )
    ctx.gpu_meta = {
        "gpu_pipeline_enabled": bool(cfg.enable),
        "gpu_pipeline_device": device,
        "gpu_pipeline_chunks": len(plans),
        "gpu_pipeline_device_index": dev.index,
        "gpu_pipeline_device_name": ctx.device_name,
    }
    return ctx


__all__ = [
    "Streams", "ChunkPlan", "PipelineConfig", "PipelineContext", "PinnedBufferPool", "InflightLimiter",
    "build_pipeline_context", "chunk_schedule", "create_streams", "record_event", "select_device", "wait_event",
]
