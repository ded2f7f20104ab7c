"""Separation backend seam - drop-in for ``audio_cut.separation.backends``.

``B200Mdx23Backend`` keeps the interface of the reference's ``MDX23OnnxBackend``
(/root/reference/src/audio_cut/separation/backends.py:90-406): ``load_model``, ``sample_rate``,
``infer_chunk(mix_chunk (N,)|(2,N) float32, stream=, non_blocking=) -> SeparationOutputs``,
``flush``, ``describe_input``, ``get_output_type``, ``reset_performance_metrics`` /
``get_performance_metrics`` - but runs pad/window -> STFT -> TFC-TDF U-Net -> iSTFT -> trim ->
stem arithmetic as sm_100a kernels through libaudiocut_b200.so.  No onnxruntime, no CPU path:
``fallback_to_cpu`` raises.
"""
from __future__ import annotations

import abc
import os
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, Optional, Union

import numpy as np
import torch

from . import _lib, ops
from .onnx_weights import load_onnx
from .unet_weights import UNetGeometry, load_npz, random_state


@dataclass
class SeparationOutputs:
    vocal: np.ndarray
    instrumental: np.ndarray


class IVocalSeparatorBackend(abc.ABC):
    @abc.abstractmethod
    def load_model(self) -> None: ...

    @abc.abstractmethod
    def sample_rate(self) -> int: ...

    @abc.abstractmethod
    def infer_chunk(self, mix_chunk: np.ndarray, **kwargs) -> SeparationOutputs: ...

    def flush(self) -> Optional[SeparationOutputs]:
        return None


def resolve_output_type(model_name: str, pref: str = "auto") -> str:
    """backends.py:198-204: 'vocal' when the file name says vocal(s) and not inst/accomp."""
    pref = (pref or "auto").strip().lower()
    if pref in ("vocal", "instrumental"):
        return pref
    name = model_name.lower()
    if ("vocal" in name) and not any(t in name for t in ("inst", "instrumental", "accomp")):
        return "vocal"
    return "instrumental"


class B200Mdx23Backend(IVocalSeparatorBackend):
    """MDX23 (Kim_Vocal geometry) separation on one B200.

    ``weights``: a ``{name: ndarray}`` state dict, a path to an ``.onnx`` (the reference's model file, read by
    ``onnx_weights.load_onnx``) or an ``.npz`` of one, or ``None`` -> the first ``*.onnx`` (else ``*.npz``) in
    ``model_dir`` (``MDX23_MODEL_FILENAME`` / ``model_filename`` select one, backends.py:144-160), or - only when
    ``allow_random_init`` - seeded random weights of the architecture.  ``n_fft`` defaults to 6144 like the reference
    (backends.py:264; env ``MDX23_N_FFT`` overrides), Kim_Vocal's native 7680 is an explicit choice.
    """

    def __init__(self, model_dir: Union[str, Path, None] = None, *, weights=None, device: str = "cuda:0",
                 precision: str = "fp16", n_fft: Optional[int] = None, align_hop: Optional[int] = None,
                 output_type: str = "auto", model_filename: Optional[str] = None, allow_random_init: bool = False,
                 geometry: Optional[UNetGeometry] = None, hop: int = 1024):
        self._model_dir = Path(model_dir) if model_dir is not None else None
        self._weights = weights
        self._device = torch.device(device)
        if self._device.type != "cuda":
            raise _lib.AudioCutError("B200Mdx23Backend runs on CUDA only (no CPU fallback)")
        # fp16 (default): IEEE-half operands on the tcgen05 tensor cores, fp32 accumulation - the 16-bit path that meets
        # the >= 40 dB stem-SDR gate; bf16: the same kernels with bfloat16 operands (fp32 range,
        # 8-bit significand: ~31 dB on the random-init network); fp32: CUDA-core FFMA path (>= 60 dB).
        if precision not in ("fp16", "bf16", "fp32"):
            raise ValueError("precision must be 'fp16', 'bf16' or 'fp32'")
        self._precision = precision
        self._dtype = {"fp16": _lib.AC_F16, "bf16": _lib.AC_BF16, "fp32": _lib.AC_F32}[precision]
        self._geo = geometry or UNetGeometry()
        self._n_fft = int(n_fft if n_fft is not None else os.getenv("MDX23_N_FFT", 6144))
        self._hop = int(hop)
        self._align_hop = int(align_hop if align_hop is not None else os.getenv("MDX23_ALIGN_HOP", 4096))
        self._output_pref = output_type
        self._model_filename = model_filename or os.getenv("MDX23_MODEL_FILENAME")
        self._allow_random = allow_random_init
        self._sr = 44100
        self._net: Optional[ops.UNet] = None
        self._model_name = "Kim_Vocal_1"
        self._resolved_output_type: Optional[str] = None
        self._perf: Dict[str, float] = {}
        self.reset_performance_metrics()

    # ---- interface -----------------------------------------------------------------------
    def sample_rate(self) -> int:
        return self._sr

    @property
    def geom(self):
        return ops.mdx_geom(self._n_fft, self._hop, self._geo.dim_f, self._geo.dim_t)

    @property
    def net(self) -> ops.UNet:
        if self._net is None:
            raise RuntimeError("B200Mdx23Backend: load_model() has not been called")
        return self._net

    @property
    def dtype(self) -> int:
        return self._dtype

    @property
    def align_hop(self) -> int:
        return self._align_hop

    def describe_input(self) -> Optional[dict]:
        return {"name": "input", "shape": [1, self._geo.dim_c, self._geo.dim_f, self._geo.dim_t]}

    def load_model(self) -> None:
        state = self._weights
        if isinstance(state, (str, Path)):
            state = self._read_weight_file(Path(state))
        if state is None and self._model_dir is not None:
            if self._model_filename:
                cand = self._model_dir / self._model_filename
                if not cand.exists():
                    raise FileNotFoundError(f"MDX23 weights not found: {cand}")
                files = [cand]
            else:
                files = sorted(self._model_dir.glob("*.onnx")) or sorted(self._model_dir.glob("*.npz"))
            if files:
                state = self._read_weight_file(files[0])
        if state is None:
            if not self._allow_random:
                where = self._model_dir if self._model_dir is not None else "(no model_dir given)"
                raise FileNotFoundError(f"no MDX23 model (.onnx / .npz) found in {where} and allow_random_init is False")
            state = random_state(self._geo)
        self._resolved_output_type = resolve_output_type(self._model_name, self._output_pref)
        idx = self._device.index if self._device.index is not None else 0
        self._net = ops.UNet(state, self._geo, device=idx)
        self.reset_performance_metrics()

    def _read_weight_file(self, path: Path):
        if not path.exists():
            raise FileNotFoundError(f"MDX23 weights not found: {path}")
        self._model_name = path.name
        if path.suffix.lower() == ".onnx":
            state, geo = load_onnx(str(path), dim_t=self._geo.dim_t)
            self._geo = geo  # the file decides g / n / l / bn / dim_f
            return state
        return load_npz(str(path))

    def reset_performance_metrics(self) -> None:
        self._perf = {"h2d_ms": 0.0, "dtoh_ms": 0.0, "compute_ms": 0.0, "chunks": 0.0, "max_alloc_bytes": 0.0}

    def get_performance_metrics(self, *, reset: bool = False) -> Dict[str, float]:
        m = dict(self._perf)
        if reset:
            self.reset_performance_metrics()
        return m

    def record_perf(self, key: str, value: float) -> None:
        if key == "max_alloc_bytes":
            self._perf[key] = max(self._perf.get(key, 0.0), float(value))
        else:
            self._perf[key] = self._perf.get(key, 0.0) + float(value)

    def get_output_type(self) -> str:
        return self._resolved_output_type or "vocal"

    def fallback_to_cpu(self) -> None:
        raise _lib.AudioCutError("audio_cut_b200 has no CPU fallback (north_star: strict GPU)")

    def infer_chunk(self, mix_chunk: np.ndarray, **kwargs) -> SeparationOutputs:
        """One pipeline chunk, host in / host out (backends.py:299-406 contract)."""
        net = self.net
        stream = kwargs.get("stream")
        mix = np.asarray(mix_chunk)
        if mix.ndim == 1:
            mix = mix[None, :]
        elif mix.ndim != 2 or mix.shape[0] != 2:
            raise ValueError("mix_chunk must have shape (N,) or (2, N)")
        n = mix.shape[1]
        if n == 0:
            z = np.zeros(0, np.float32)
            return SeparationOutputs(vocal=z, instrumental=z.copy())
        host = torch.from_numpy(np.ascontiguousarray(mix, dtype=np.float32))
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        with torch.cuda.device(self._device), torch.cuda.stream(stream if stream is not None else torch.cuda.current_stream()):
            ev[0].record()
            dev = host.pin_memory().to(self._device, non_blocking=True)
            ev[1].record()
            v, i, _ = ops.separate_track(net, dev, [(0, n, 0, n)], self.geom, align_hop=self._align_hop,
                                         output_is_vocal=self.get_output_type() == "vocal", dtype=self._dtype)
            ev[2].record()
            out = torch.stack([v, i]).cpu()
            ev[3].record()
            ev[3].synchronize()
        if _lib.load().ac_debug_tc_aborted() != 0:
            raise _lib.AudioCutError("a tcgen05 kernel hit its mbarrier watchdog during this chunk: separation output is invalid")
        self.record_perf("h2d_ms", ev[0].elapsed_time(ev[1]))
        self.record_perf("compute_ms", ev[1].elapsed_time(ev[2]))
        self.record_perf("dtoh_ms", ev[2].elapsed_time(ev[3]))
        self.record_perf("max_alloc_bytes", torch.cuda.max_memory_allocated(self._device))
        self.record_perf("chunks", 1.0)
        out = out.numpy()
        return SeparationOutputs(vocal=out[0].copy(), instrumental=out[1].copy())


__all__ = ["SeparationOutputs", "IVocalSeparatorBackend", "B200Mdx23Backend", "resolve_output_type"]
