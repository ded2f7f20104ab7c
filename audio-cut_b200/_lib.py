"""ctypes binding of ``libaudiocut_b200.so`` (the C ABI declared in include/audiocut_b200.h).

There is no fallback: if the shared library is missing, or the machine has no sm_100
device, importing callers get a loud ``RuntimeError`` - never a silent CPU path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libaudiocut_b200.so")

AC_F32, AC_BF16, AC_F16 = 0, 1, 2

EXPORTS = [
    "ac_init", "ac_last_error", "ac_abi_version", "ac_launch_count", "ac_frame_count", "ac_frame_rms", "ac_frame_rms_segments",
    "ac_stft_mdx", "ac_istft_mdx", "ac_unet_create", "ac_unet_destroy", "ac_unet_param_floats",
    "ac_unet_workspace_bytes", "ac_unet_forward", "ac_unet_set_debug", "ac_track_window_count",
    "ac_track_workspace_bytes", "ac_separate_track", "ac_separate_track_ex", "ac_separate_track_pipelined", "ac_stft_features_workspace_bytes", "ac_stft_features",
    "ac_zero_crossing_rate", "ac_debug_tc_aborted", "ac_profile_begin", "ac_profile_collect",
    "ac_tempogram_stats", "ac_host_beat_dp", "ac_downmix_mono", "ac_track_stats", "ac_debug_conv3x3", "ac_debug_conv3x3_chain", "ac_pyin_frame_count", "ac_pyin_workspace_bytes", "ac_pyin",
    "ac_lpc_frame_count", "ac_lpc_formants", "ac_refine_cut_points", "ac_quiet_lookup_db",
    "ac_host_is_pinned", "ac_copy_h2d_async",
    "ac_pcm_decode", "ac_peak_normalize", "ac_pcm_pack", "ac_resample_out_len", "ac_resample_poly",
]


class MdxGeom(C.Structure):
    _fields_ = [("n_fft", C.c_int), ("hop", C.c_int), ("dim_f", C.c_int), ("dim_t", C.c_int)]


class UNetGeom(C.Structure):
    _fields_ = [("dim_f", C.c_int), ("dim_t", C.c_int), ("dim_c", C.c_int), ("g", C.c_int), ("n", C.c_int),
                ("l", C.c_int), ("bn", C.c_int)]


class ChunkDesc(C.Structure):
    _fields_ = [("chunk_start", C.c_longlong), ("eff_start", C.c_longlong), ("eff_end", C.c_longlong),
                ("chunk_len", C.c_int), ("reserved", C.c_int)]


class TrackParams(C.Structure):
    _fields_ = [("mdx", MdxGeom), ("align_hop", C.c_int), ("n_channels", C.c_int), ("output_is_vocal", C.c_int),
                ("dtype", C.c_int), ("max_batch", C.c_int), ("reserved", C.c_int)]


class KernelStat(C.Structure):
    _fields_ = [("name", C.c_char_p), ("launches", C.c_longlong), ("total_ms", C.c_double), ("flops", C.c_double),
                ("bytes", C.c_double)]


class FeatSegment(C.Structure):
    _fields_ = [("start", C.c_longlong), ("len", C.c_longlong), ("frame_off", C.c_longlong)]


_lib: Optional[C.CDLL] = None
_inited_devices = set()


def load() -> C.CDLL:
    """dlopen the library and declare every prototype (no CUDA call is made)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C audio-cut_b200/csrc). audio_cut_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    vp, ll, i, sz = C.c_void_p, C.c_longlong, C.c_int, C.c_size_t
    lib.ac_init.argtypes, lib.ac_init.restype = [i], i
    lib.ac_last_error.argtypes, lib.ac_last_error.restype = [], C.c_char_p
    lib.ac_abi_version.argtypes, lib.ac_abi_version.restype = [], i
    lib.ac_launch_count.argtypes, lib.ac_launch_count.restype = [], ll
    lib.ac_frame_count.argtypes, lib.ac_frame_count.restype = [ll, i, i, i], ll
    lib.ac_frame_rms.argtypes, lib.ac_frame_rms.restype = [vp, ll, i, i, i, vp, vp], i
    lib.ac_frame_rms_segments.argtypes = [vp, ll, C.POINTER(FeatSegment), i, i, i, i, vp, vp]
    lib.ac_frame_rms_segments.restype = i
    lib.ac_zero_crossing_rate.argtypes, lib.ac_zero_crossing_rate.restype = [vp, ll, i, i, vp, vp], i
    lib.ac_stft_mdx.argtypes, lib.ac_stft_mdx.restype = [vp, vp, i, C.POINTER(MdxGeom), i, vp], i
    lib.ac_istft_mdx.argtypes, lib.ac_istft_mdx.restype = [vp, vp, i, C.POINTER(MdxGeom), i, vp], i
    lib.ac_unet_create.argtypes = [C.POINTER(UNetGeom), vp, sz, C.POINTER(vp)]
    lib.ac_unet_create.restype = i
    lib.ac_unet_destroy.argtypes, lib.ac_unet_destroy.restype = [vp], None
    lib.ac_unet_param_floats.argtypes, lib.ac_unet_param_floats.restype = [C.POINTER(UNetGeom)], sz
    lib.ac_unet_workspace_bytes.argtypes, lib.ac_unet_workspace_bytes.restype = [vp, i, i], sz
    lib.ac_unet_forward.argtypes, lib.ac_unet_forward.restype = [vp, vp, vp, i, i, vp, sz, vp], i
    lib.ac_unet_set_debug.argtypes, lib.ac_unet_set_debug.restype = [vp, i], i
    lib.ac_debug_tc_aborted.argtypes, lib.ac_debug_tc_aborted.restype = [], i
    lib.ac_debug_conv3x3.argtypes = [vp, vp, i, i, i, i, vp, vp, vp, i, i, C.POINTER(C.c_float), vp]
    lib.ac_debug_conv3x3.restype = i
    lib.ac_debug_conv3x3_chain.argtypes = [vp, vp, vp, i, i, i, i, vp, vp, vp, i, i, C.POINTER(C.c_float), vp]
    lib.ac_debug_conv3x3_chain.restype = i
    lib.ac_track_window_count.argtypes = [C.POINTER(ChunkDesc), i, C.POINTER(TrackParams)]
    lib.ac_track_window_count.restype = i
    lib.ac_track_workspace_bytes.argtypes = [vp, C.POINTER(ChunkDesc), i, C.POINTER(TrackParams)]
    lib.ac_track_workspace_bytes.restype = sz
    lib.ac_separate_track.argtypes = [vp, vp, ll, C.POINTER(ChunkDesc), i, C.POINTER(TrackParams), vp, vp, vp, vp, sz, vp]
    lib.ac_separate_track.restype = i
    lib.ac_separate_track_ex.argtypes = [vp, vp, ll, C.POINTER(ChunkDesc), i, C.POINTER(TrackParams), vp, vp, vp, vp, vp, sz, vp]
    lib.ac_separate_track_ex.restype = i
    lib.ac_separate_track_pipelined.argtypes = [vp, vp, vp, ll, C.POINTER(ChunkDesc), i, C.POINTER(TrackParams), vp, vp, vp, vp, vp, vp,
                                                vp, sz, vp, vp, vp]
    lib.ac_separate_track_pipelined.restype = i
    lib.ac_stft_features_workspace_bytes.argtypes = [C.POINTER(FeatSegment), i, i]
    lib.ac_stft_features_workspace_bytes.restype = sz
    lib.ac_stft_features.argtypes = [vp, C.POINTER(FeatSegment), i, i, i, vp, vp, vp, vp, vp, vp, sz, vp]
    lib.ac_stft_features.restype = i
    lib.ac_tempogram_stats.argtypes, lib.ac_tempogram_stats.restype = [vp, ll, i, vp, vp, vp, vp], i
    lib.ac_host_beat_dp.argtypes, lib.ac_host_beat_dp.restype = [vp, i, i, C.c_float, vp, vp], i
    lib.ac_downmix_mono.argtypes, lib.ac_downmix_mono.restype = [vp, i, ll, vp, vp], i
    lib.ac_track_stats.argtypes, lib.ac_track_stats.restype = [vp, vp, vp, ll, vp, vp], i
    lib.ac_pyin_frame_count.argtypes, lib.ac_pyin_frame_count.restype = [ll, i], ll
    lib.ac_pyin_workspace_bytes.argtypes, lib.ac_pyin_workspace_bytes.restype = [ll, i, i, C.c_float, C.c_float], sz
    lib.ac_pyin.argtypes, lib.ac_pyin.restype = [vp, ll, i, i, C.c_float, C.c_float, vp, vp, vp, vp, sz, vp], i
    lib.ac_lpc_frame_count.argtypes, lib.ac_lpc_frame_count.restype = [ll, i, i], ll
    lib.ac_lpc_formants.argtypes, lib.ac_lpc_formants.restype = [vp, ll, i, i, i, vp, vp, vp], i
    d = C.c_double
    lib.ac_refine_cut_points.argtypes = [vp, vp, ll, i, vp, i, i, i, i, d, d, i, i, i, vp, vp, vp]
    lib.ac_refine_cut_points.restype = i
    lib.ac_quiet_lookup_db.argtypes, lib.ac_quiet_lookup_db.restype = [vp, ll, i, vp, vp], i
    lib.ac_host_is_pinned.argtypes, lib.ac_host_is_pinned.restype = [vp, sz], i
    lib.ac_copy_h2d_async.argtypes, lib.ac_copy_h2d_async.restype = [vp, vp, sz, vp], i
    lib.ac_pcm_decode.argtypes, lib.ac_pcm_decode.restype = [vp, ll, i, i, i, vp, vp], i
    lib.ac_peak_normalize.argtypes, lib.ac_peak_normalize.restype = [vp, ll, vp, vp], i
    lib.ac_pcm_pack.argtypes, lib.ac_pcm_pack.restype = [vp, ll, i, i, vp, vp], i
    lib.ac_resample_out_len.argtypes, lib.ac_resample_out_len.restype = [ll, i, i], ll
    lib.ac_resample_poly.argtypes = [vp, C.POINTER(ll), C.POINTER(ll), i, i, i, vp, i, i, i, ll, vp, vp]
    lib.ac_resample_poly.restype = i
    lib.ac_profile_begin.argtypes, lib.ac_profile_begin.restype = [], i
    lib.ac_profile_collect.argtypes, lib.ac_profile_collect.restype = [C.POINTER(KernelStat), i], i
    _lib = lib
    return lib


class AudioCutError(RuntimeError):
    pass


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().ac_last_error().decode("utf-8", "replace")
        raise AudioCutError(f"{what or 'libaudiocut_b200'} failed (rc={rc}): {msg}")


def init(device_index: int = 0) -> C.CDLL:
    """Load the library and bring up ``cuda:device_index``; raises when there is no B200."""
    lib = load()
    if device_index not in _inited_devices:
        check(lib.ac_init(int(device_index)), "ac_init")
        _inited_devices.add(device_index)
    return lib


def profile_begin() -> None:
    check(load().ac_profile_begin(), "ac_profile_begin")


def profile_collect():
    """[{name, launches, total_ms, flops, bytes}] for every kernel class that launched."""
    arr = (KernelStat * 32)()
    n = load().ac_profile_collect(arr, 32)
    if n < 0:
        check(n, "ac_profile_collect")
    return [
        {"name": arr[k].name.decode(), "launches": int(arr[k].launches), "total_ms": float(arr[k].total_ms),
         "flops": float(arr[k].flops), "bytes": float(arr[k].bytes)}
        for k in range(n) if arr[k].launches
    ]


def ptr(t) -> int:
    """Device pointer of a torch tensor (or None -> NULL)."""
    return 0 if t is None else int(t.data_ptr())


def stream_ptr(stream=None) -> int:
    import torch

    s = stream if stream is not None else torch.cuda.current_stream()
    return int(s.cuda_stream)
