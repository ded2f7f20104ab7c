"""PCM in / PCM out on the GPU - the data formats either side of the hot path.

* ``decode_pcm`` + ``load_wav``: what ``AudioProcessor.load_audio`` gets out of ``librosa.load(path, sr, mono=True)`` for a
  16- / 24-bit PCM WAV whose rate already is the target rate (/root/reference/src/vocal_smart_splitter/utils/
  audio_processor.py:32-60): libsndfile's float normalisation, channel mean, then ``audio / max|audio|``.  The bytes go to the
  device as they lie in the file; the float32 track never exists on the host.
* ``pack_pcm`` + ``write_wav``: ``export_audio(..., "wav")`` = ``sf.write(path, audio, sr, subtype="PCM_24")``
  (utils/audio_export.py:70-112); ``format="int16"`` is the sample conversion of the MP3 writer (:113-133).

Container parsing / writing is Python's ``wave`` module (RIFF headers are not bandwidth work).
"""
from __future__ import annotations

import wave
from typing import Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr

PCM_FORMATS = {"PCM_24": 0, "PCM_24_clip": 1, "int16": 2}


def decode_pcm(raw: torch.Tensor, n_frames: int, channels: int, bits: int, *, mono: bool = True, normalize: bool = False) -> torch.Tensor:
    """raw: uint8 CUDA tensor of interleaved little-endian PCM frames -> float32 [n_frames] (mono) or [channels, n_frames]."""
    if not raw.is_cuda or raw.dtype != torch.uint8:
        raise _lib.AudioCutError("decode_pcm needs a uint8 CUDA tensor (there is no CPU fallback)")
    lib = _lib.init(raw.device.index or 0)
    raw = raw.contiguous()
    if raw.numel() < n_frames * channels * (bits // 8):
        raise ValueError("raw buffer shorter than n_frames * channels * bytes per sample")
    out = torch.empty(n_frames if mono else (channels, n_frames), dtype=torch.float32, device=raw.device)
    check(lib.ac_pcm_decode(ptr(raw), n_frames, channels, bits, int(mono), ptr(out), stream_ptr()), "ac_pcm_decode")
    if normalize:
        peak_normalize_(out)
    return out


def peak_normalize_(x: torch.Tensor) -> torch.Tensor:
    """In place ``x / max|x|`` (skipped when the signal is all zero), audio_processor.py:54-56."""
    lib = _lib.init(x.device.index or 0)
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
    scratch = torch.empty(1, dtype=torch.int32, device=x.device)
    check(lib.ac_peak_normalize(ptr(x), x.numel(), ptr(scratch), stream_ptr()), "ac_peak_normalize")
    return x


def load_wav(path: str, *, device: str = "cuda", mono: bool = True, normalize: bool = True) -> Tuple[torch.Tensor, int]:
    """(audio on the device, sample rate) of a 16- / 24-bit PCM WAV file."""
    with wave.open(path, "rb") as w:
        ch, width, sr, n = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
        if width not in (2, 3):
            raise ValueError(f"{path}: only 16- and 24-bit PCM are decoded on the GPU (sample width {width})")
        data = w.readframes(n)
    host = torch.frombuffer(bytearray(data), dtype=torch.uint8)
    raw = host.pin_memory().to(device, non_blocking=True)
    return decode_pcm(raw, n, ch, 8 * width, mono=mono, normalize=normalize), sr


def pack_pcm(x: torch.Tensor, format: str = "PCM_24") -> torch.Tensor:
    """x: float32 CUDA tensor [n] or [channels, n] -> uint8 tensor of interleaved frames."""
    if not x.is_cuda:
        raise _lib.AudioCutError("pack_pcm needs a CUDA tensor (there is no CPU fallback)")
    lib = _lib.init(x.device.index or 0)
    x = x.contiguous().float()
    ch, n = (1, x.numel()) if x.dim() == 1 else (int(x.shape[0]), int(x.shape[1]))
    fmt = PCM_FORMATS[format]
    out = torch.empty(n * ch * (2 if fmt == 2 else 3), dtype=torch.uint8, device=x.device)
    check(lib.ac_pcm_pack(ptr(x), n, ch, fmt, ptr(out), stream_ptr()), "ac_pcm_pack")
    return out


def write_wav(path: str, x: torch.Tensor, sample_rate: int, subtype: str = "PCM_24") -> None:
    """``sf.write(path, audio, sample_rate, subtype=subtype)`` for a device-resident stem / segment."""
    if subtype not in ("PCM_24", "PCM_16"):
        raise ValueError("subtype must be PCM_24 or PCM_16")
    data = pack_pcm(x, "PCM_24" if subtype == "PCM_24" else "int16").cpu().numpy().tobytes()
    ch = 1 if x.dim() == 1 else int(x.shape[0])
    with wave.open(path, "wb") as w:
        w.setnchannels(ch)
        w.setsampwidth(3 if subtype == "PCM_24" else 2)
        w.setframerate(int(sample_rate))
        w.writeframes(data)


__all__ = ["decode_pcm", "peak_normalize_", "load_wav", "pack_pcm", "write_wav", "PCM_FORMATS"]
