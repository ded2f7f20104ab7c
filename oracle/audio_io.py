"""Oracle for the PCM formats either side of the path and the VAD resampler - test infrastructure.

The reference delegates all of it to third-party libraries that are not in this image, so their published arithmetic is
restated (PARITY UNPINNED against libsndfile itself; pinned against Python's own ``wave`` / ``struct`` for the container and
the integer layout, and against scipy - which IS installed - for the resampler):

* ``sf.write(path, audio, sr, subtype="PCM_24")``   /root/reference/src/vocal_smart_splitter/utils/audio_export.py:109-112
      libsndfile 1.x ``pcm.c``: ``f2let_array``  value = lrintf(x * 0x7FFFFF), bytes value, value>>8, value>>16 (no clipping
      by default); ``f2let_clip_array`` (after SFC_SET_CLIPPING) scales by 2^31, saturates and keeps the top three bytes.
* the MP3 writer's int16 conversion                  audio_export.py:120-123   np.round(clip(x,-1,1) * 32767).astype(int16)
* ``librosa.load(path, sr, mono=True)`` + peak norm  utils/audio_processor.py:45-56   sf_read_float: PCM_16 / 2^15,
      PCM_24 / 2^23; librosa.to_mono = mean over channels; audio / max|audio|.
* ``librosa.resample(audio, orig_sr=44100, target_sr=16000)`` per VAD chunk   core/vocal_pause_detector.py:190 - librosa's
      default ``soxr_hq`` needs libsoxr (absent); ``res_type="polyphase"`` is ``scipy.signal.resample_poly``, used here as is.
"""
from __future__ import annotations

import numpy as np


def pcm24_bytes(x: np.ndarray, clip: bool = False) -> bytes:
    """x: float32 [n] or [n, ch] (frames first, as soundfile takes it) -> interleaved little-endian 24-bit frames."""
    a = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
    if clip:
        sc = a * np.float32(2147483648.0)
        v = np.rint(np.clip(sc.astype(np.float64), -2147483648.0, 2147483647.0)).astype(np.int64)
        v = np.where(sc >= np.float32(2147483648.0), 0x7FFFFFFF, v)
        v = np.where(sc <= np.float32(-2147483648.0), -0x80000000, v) >> 8
    else:
        prod = (a * np.float32(8388607.0)).astype(np.float32)  # float32 product, rounded to nearest even like lrintf
        v = np.rint(prod.astype(np.float64)).astype(np.int64)
    v = v & 0xFFFFFF
    out = np.empty((a.size, 3), np.uint8)
    out[:, 0], out[:, 1], out[:, 2] = v & 0xFF, (v >> 8) & 0xFF, (v >> 16) & 0xFF
    return out.tobytes()


def int16_bytes(x: np.ndarray) -> bytes:
    a = np.clip(np.ascontiguousarray(x, dtype=np.float32).reshape(-1), -1.0, 1.0)
    return np.round(a * np.float32(32767.0)).astype("<i2").tobytes()


def decode_pcm(data: bytes, channels: int, bits: int, mono: bool = True, normalize: bool = False) -> np.ndarray:
    """Interleaved PCM bytes -> float32 [n] (mono) or [ch, n]."""
    raw = np.frombuffer(data, dtype=np.uint8)
    if bits == 16:
        v = raw.view("<i2").astype(np.float32) * np.float32(1.0 / 32768.0)
    elif bits == 24:
        b = raw.reshape(-1, 3).astype(np.int32)
        i = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        i = np.where(i & 0x800000, i - (1 << 24), i)
        v = i.astype(np.float32) * np.float32(1.0 / 8388608.0)
    else:
        raise ValueError("bits")
    v = v.reshape(-1, channels).T
    out = (np.mean(v, axis=0) if channels > 1 else v[0]) if mono else v
    out = np.ascontiguousarray(out, dtype=np.float32)
    if normalize and np.max(np.abs(out)) > 0:
        out = out / np.max(np.abs(out))
    return out


def resample(x: np.ndarray, sr_in: int = 44100, sr_out: int = 16000) -> np.ndarray:
    import scipy.signal

    return scipy.signal.resample_poly(np.asarray(x, dtype=np.float32), sr_out, sr_in).astype(np.float32)
