"""Thin torch-tensor wrappers over the C ABI (device memory and streams are torch's; the math is ours)."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import AC_BF16, AC_F16, AC_F32, ChunkDesc, FeatSegment, MdxGeom, TrackParams, UNetGeom, check, ptr, stream_ptr
from .unet_weights import UNetGeometry, pack_blob


def _dev_index(t: torch.Tensor) -> int:
    if not t.is_cuda:
        raise _lib.AudioCutError("audio_cut_b200 ops need CUDA tensors (there is no CPU fallback)")
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def _torch_dtype(dtype: int):
    return {AC_F32: torch.float32, AC_BF16: torch.bfloat16, AC_F16: torch.float16}[dtype]


def _ac_dtype(t: torch.Tensor) -> int:
    try:
        return {torch.float32: AC_F32, torch.bfloat16: AC_BF16, torch.float16: AC_F16}[t.dtype]
    except KeyError:
        raise _lib.AudioCutError(f"unsupported tensor dtype {t.dtype} (float32, float16 or bfloat16)") from None


def frame_count(n: int, frame: int, hop: int, center: bool = True) -> int:
    return int(_lib.load().ac_frame_count(int(n), int(frame), int(hop), int(bool(center))))


def frame_rms(x: torch.Tensor, frame_length: int, hop_length: int, center: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """librosa.feature.rms(y, frame_length, hop_length)[0] on the GPU (``out``: preallocated float32 slice)."""
    lib = _lib.init(_dev_index(x))
    x = x.contiguous().float()
    n = x.numel()
    nf = frame_count(n, frame_length, hop_length, center)
    if out is None:
        out = torch.empty(nf, dtype=torch.float32, device=x.device)
    assert out.numel() == nf and out.is_contiguous() and out.dtype == torch.float32
    if out.numel():
        check(lib.ac_frame_rms(ptr(x), n, frame_length, hop_length, int(center), ptr(out), stream_ptr()), "ac_frame_rms")
    return out


def frame_rms_segments(x: torch.Tensor, segments: Sequence[Tuple[int, int, int]], frame_length: int, hop_length: int, *,
                       total_frames: int, center: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """frame_rms of many slices of ``x`` in ONE launch; segments = (start, len, frame_off) as for ``stft_features``;
    slice i's ``frame_count(len, ...)`` values land at ``out[frame_off:]``."""
    lib = _lib.init(_dev_index(x))
    x = x.contiguous().float()
    if out is None:
        out = torch.empty(total_frames, dtype=torch.float32, device=x.device)
    assert out.numel() >= total_frames and out.is_contiguous() and out.dtype == torch.float32
    if segments:
        segs = (FeatSegment * len(segments))(*[FeatSegment(int(s), int(l), int(o)) for s, l, o in segments])
        check(lib.ac_frame_rms_segments(ptr(x), x.numel(), segs, len(segments), frame_length, hop_length, int(center), ptr(out),
                                        stream_ptr()), "ac_frame_rms_segments")
    return out


def zero_crossing_rate(x: torch.Tensor, frame_length: int = 2048, hop_length: int = 512) -> torch.Tensor:
    lib = _lib.init(_dev_index(x))
    x = x.contiguous().float()
    n = x.numel()
    out = torch.empty(frame_count(n, frame_length, hop_length, True), dtype=torch.float32, device=x.device)
    if out.numel():
        check(lib.ac_zero_crossing_rate(ptr(x), n, frame_length, hop_length, ptr(out), stream_ptr()), "ac_zero_crossing_rate")
    return out


def mdx_geom(n_fft=7680, hop=1024, dim_f=3072, dim_t=256) -> MdxGeom:
    return MdxGeom(int(n_fft), int(hop), int(dim_f), int(dim_t))


def stft_mdx(wave: torch.Tensor, geom: MdxGeom, dtype: int = AC_F32) -> torch.Tensor:
    """wave [B,2,W] f32 -> spec [B,dim_t,dim_f,4] ("TFC" layout, see include/audiocut_b200.h)."""
    lib = _lib.init(_dev_index(wave))
    W = geom.hop * (geom.dim_t - 1)
    assert wave.dim() == 3 and wave.shape[1] == 2 and wave.shape[2] == W, wave.shape
    wave = wave.contiguous().float()
    B = wave.shape[0]
    spec = torch.empty((B, geom.dim_t, geom.dim_f, 4), dtype=_torch_dtype(dtype), device=wave.device)
    check(lib.ac_stft_mdx(ptr(wave), ptr(spec), B, C.byref(geom), dtype, stream_ptr()), "ac_stft_mdx")
    return spec


def istft_mdx(spec: torch.Tensor, geom: MdxGeom) -> torch.Tensor:
    lib = _lib.init(_dev_index(spec))
    dtype = _ac_dtype(spec)
    spec = spec.contiguous()
    B = spec.shape[0]
    assert tuple(spec.shape[1:]) == (geom.dim_t, geom.dim_f, 4), spec.shape
    W = geom.hop * (geom.dim_t - 1)
    wave = torch.empty((B, 2, W), dtype=torch.float32, device=spec.device)
    check(lib.ac_istft_mdx(ptr(spec), ptr(wave), B, C.byref(geom), dtype, stream_ptr()), "ac_istft_mdx")
    return wave


def onnx_to_tfc(x: torch.Tensor) -> torch.Tensor:
    """[B,4,F,T] (ONNX tensor of backends.py:356) -> [B,T,F,4]."""
    return x.permute(0, 3, 2, 1).contiguous()


def tfc_to_onnx(x: torch.Tensor) -> torch.Tensor:
    return x.permute(0, 3, 2, 1).contiguous()


class UNet:
    """Device-resident TFC-TDF U-Net (replaces the onnxruntime session of backends.py:216-253, 358)."""

    def __init__(self, state: Dict[str, np.ndarray], geo: UNetGeometry = UNetGeometry(), device: int = 0):
        self.geo = geo
        self.device = torch.device("cuda", device)
        self._lib = _lib.init(device)
        g = UNetGeom(geo.dim_f, geo.dim_t, geo.dim_c, geo.g, geo.n, geo.l, geo.bn)
        self._g = g
        blob = pack_blob(state, geo)
        want = int(self._lib.ac_unet_param_floats(C.byref(g)))
        if blob.size != want:
            raise _lib.AudioCutError(f"parameter blob has {blob.size} floats, geometry needs {want}")
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self._lib.ac_unet_create(C.byref(g), blob.ctypes.data_as(C.c_void_p), blob.size, C.byref(h)), "ac_unet_create")
        self.handle = h
        self._ws: Optional[torch.Tensor] = None

    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            try:
                self._lib.ac_unet_destroy(h)
            except Exception:
                pass
            self.handle = None

    def set_debug(self, force_simt: int) -> None:
        check(self._lib.ac_unet_set_debug(self.handle, int(force_simt)))

    def workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self._ws

    def forward(self, spec: torch.Tensor) -> torch.Tensor:
        """spec [B,dim_t,dim_f,4] float32, float16 or bfloat16 (TFC layout) -> same shape/dtype."""
        dtype = _ac_dtype(spec)
        spec = spec.contiguous()
        B = spec.shape[0]
        assert tuple(spec.shape[1:]) == (self.geo.dim_t, self.geo.dim_f, 4), spec.shape
        out = torch.empty_like(spec)
        nbytes = int(self._lib.ac_unet_workspace_bytes(self.handle, B, dtype))
        ws = self.workspace(nbytes)
        check(self._lib.ac_unet_forward(self.handle, ptr(spec), ptr(out), B, dtype, ptr(ws), ws.numel(), stream_ptr()), "ac_unet_forward")
        return out


def debug_conv3x3(x: torch.Tensor, w: np.ndarray, scale: torch.Tensor, shift: torch.Tensor, impl: int, iters: int = 1):
    """One bf16 3x3 conv layer (test hook): x [B,T,F,C] bf16, w [C,C,3,3] float32 -> (y, mean ms per launch).

    impl 0 = CUDA cores, 1 = streaming tcgen05, 2 = weight-stationary tcgen05; the tensor-core kernels
    work on the CG8 layout [B,T,C/8,F,8], the conversion to and from channels-last is done here."""
    lib = _lib.init(_dev_index(x))
    assert x.dtype in (torch.bfloat16, torch.float16) and x.dim() == 4
    if x.dtype == torch.float16:
        impl |= 16  # IEEE-half operands
    B, T, F, Cc = x.shape
    if impl & 15:
        x = x.view(B, T, F, Cc // 8, 8).permute(0, 1, 3, 2, 4)
    x = x.contiguous()
    w = np.ascontiguousarray(w, dtype=np.float32)
    assert w.shape == (Cc, Cc, 3, 3)
    y = torch.empty((B, T, F, Cc), dtype=x.dtype, device=x.device)
    ms = C.c_float(0.0)
    check(lib.ac_debug_conv3x3(ptr(x), ptr(y), B, T, F, Cc, w.ctypes.data_as(C.c_void_p), ptr(scale.float().contiguous()),
                               ptr(shift.float().contiguous()), impl, iters, C.byref(ms), stream_ptr()), "ac_debug_conv3x3")
    if impl & 15:
        y = y.view(B, T, Cc // 8, F, 8).permute(0, 1, 3, 2, 4).contiguous().view(B, T, F, Cc)
    return y, float(ms.value)


def debug_conv3x3_chain(x: torch.Tensor, w: np.ndarray, scale: torch.Tensor, shift: torch.Tensor, impl: int, iters: int = 1):
    """Three chained 16-bit 3x3 conv layers (test hook, C = 48): x [B,T,F,C], w [3,C,C,3,3] float32, scale / shift [3,C]
    -> (y, mean ms per pass).  impl 0 = three weight-stationary launches, 1 = the fused kernel."""
    lib = _lib.init(_dev_index(x))
    assert x.dtype in (torch.bfloat16, torch.float16) and x.dim() == 4
    B, T, F, Cc = x.shape
    x = x.view(B, T, F, Cc // 8, 8).permute(0, 1, 3, 2, 4).contiguous()
    w = np.ascontiguousarray(w, dtype=np.float32)
    assert w.shape == (3, Cc, Cc, 3, 3)
    y = torch.empty_like(x)
    tmp = torch.empty_like(x)
    ms = C.c_float(0.0)
    sc = scale.float().contiguous()
    sh = shift.float().contiguous()
    assert tuple(sc.shape) == (3, Cc) and tuple(sh.shape) == (3, Cc)
    check(lib.ac_debug_conv3x3_chain(ptr(x), ptr(y), ptr(tmp), B, T, F, Cc, w.ctypes.data_as(C.c_void_p), ptr(sc), ptr(sh),
                                     impl | (16 if x.dtype == torch.float16 else 0), iters, C.byref(ms), stream_ptr()),
          "ac_debug_conv3x3_chain")
    y = y.view(B, T, Cc // 8, F, 8).permute(0, 1, 3, 2, 4).contiguous().view(B, T, F, Cc)
    return y, float(ms.value)


def make_chunk_descs(bounds: Sequence[Tuple[int, int, int, int]]):
    """bounds: (chunk_start, chunk_end, eff_start, eff_end) per chunk, in samples."""
    arr = (ChunkDesc * max(1, len(bounds)))()
    for i, (cs, ce, es, ee) in enumerate(bounds):
        arr[i] = ChunkDesc(int(cs), int(es), int(ee), int(ce - cs), 0)
    return arr


def separate_track(net: UNet, mix: torch.Tensor, bounds: Sequence[Tuple[int, int, int, int]], geom: MdxGeom, *,
                   align_hop: int = 4096, output_is_vocal: bool = True, dtype: int = AC_F32, max_batch: int = 0, out=None,
                   chunk_vocal: Optional[torch.Tensor] = None, host_mix: int = 0, host_out: Optional[Tuple[int, int]] = None,
                   copy_stream: Optional[torch.cuda.Stream] = None, uploaded_event: Optional[torch.cuda.Event] = None):
    """mix [n_ch, N] f32 on the GPU -> (vocal [N], instrumental [N], weight [N]); ``out`` = preallocated triple.
    ``chunk_vocal`` (optional, float32 [sum of chunk lengths]): receives every chunk's own pre-stitch vocal output.
    With ``copy_stream``: the copies are pipelined against the window batches (ac_separate_track_pipelined) - ``host_mix`` =
    address of the page-locked [n_ch, N] float32 source that is uploaded INTO ``mix`` piece by piece (0: ``mix`` is already
    resident), ``host_out`` = addresses of two page-locked [N] float32 arrays receiving the stems as they become final,
    ``uploaded_event`` is recorded on ``copy_stream`` when the whole mix is on the device."""
    assert mix.dim() == 2 and mix.shape[0] in (1, 2)
    mix = mix.contiguous().float()
    n = mix.shape[1]
    descs = make_chunk_descs(bounds)
    tp = TrackParams(geom, int(align_hop), int(mix.shape[0]), int(bool(output_is_vocal)), int(dtype), int(max_batch), 0)
    lib = net._lib
    nbytes = int(lib.ac_track_workspace_bytes(net.handle, descs, len(bounds), C.byref(tp)))
    ws = net.workspace(nbytes)
    if out is not None:
        vocal, instr, weight = out
        for t in out:
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == n
    else:
        vocal = torch.empty(n, dtype=torch.float32, device=mix.device)
        instr = torch.empty_like(vocal)
        weight = torch.empty_like(vocal)
    if chunk_vocal is not None:
        need = sum(max(0, ce - cs) for cs, ce, _, _ in bounds)
        assert chunk_vocal.is_cuda and chunk_vocal.dtype == torch.float32 and chunk_vocal.is_contiguous() and chunk_vocal.numel() >= need
    if copy_stream is not None:
        hv, hi = host_out if host_out is not None else (0, 0)
        if uploaded_event is not None and not uploaded_event.cuda_event:  # torch creates the CUDA event lazily
            uploaded_event.record(copy_stream)
        check(lib.ac_separate_track_pipelined(net.handle, ptr(mix), C.c_void_p(host_mix or None), n, descs, len(bounds), C.byref(tp),
                                              ptr(vocal), ptr(instr), ptr(weight), ptr(chunk_vocal), C.c_void_p(hv or None),
                                              C.c_void_p(hi or None), ptr(ws), ws.numel(), stream_ptr(),
                                              C.c_void_p(copy_stream.cuda_stream),
                                              C.c_void_p(uploaded_event.cuda_event if uploaded_event is not None else None)),
              "ac_separate_track_pipelined")
        return vocal, instr, weight
    check(lib.ac_separate_track_ex(net.handle, ptr(mix), n, descs, len(bounds), C.byref(tp), ptr(vocal), ptr(instr), ptr(weight),
                                   ptr(chunk_vocal), ptr(ws), ws.numel(), stream_ptr()), "ac_separate_track_ex")
    return vocal, instr, weight


def stft_features(x: torch.Tensor, segments: Sequence[Tuple[int, int, int]], hop: int, sr: int, *, total_frames: int,
                  want=("flatness", "onset_mean")) -> Dict[str, torch.Tensor]:
    """segments: (start, len, frame_off).  Returns the requested series, each ``total_frames`` long."""
    lib = _lib.init(_dev_index(x))
    x = x.contiguous().float()
    segs = (FeatSegment * len(segments))(*[FeatSegment(int(s), int(l), int(o)) for s, l, o in segments])
    names = ("flatness", "onset_mean", "onset_median", "centroid", "low_ratio")
    out = {k: torch.zeros(total_frames, dtype=torch.float32, device=x.device) for k in names if k in want}
    nbytes = int(lib.ac_stft_features_workspace_bytes(segs, len(segments), hop))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    check(lib.ac_stft_features(ptr(x), segs, len(segments), hop, sr, *[ptr(out.get(k)) for k in names], ptr(ws), nbytes,
                               stream_ptr()), "ac_stft_features")
    return out


def tempo_prior(win: int, sr: int, hop: int, start_bpm: float = 120.0, std_bpm: float = 1.0, max_tempo: float = 320.0):
    """(bpms[win], logprior[win]) of librosa.feature.rhythm.tempo (lag 0 -> inf bpm, excluded)."""
    bpms = np.empty(win, dtype=np.float64)
    bpms[0] = np.inf
    bpms[1:] = 60.0 * sr / (hop * np.arange(1.0, win))
    with np.errstate(divide="ignore", invalid="ignore"):
        logprior = -0.5 * ((np.log2(bpms) - np.log2(start_bpm)) / std_bpm) ** 2
    logprior[: int(np.argmax(bpms < max_tempo))] = -np.inf
    return bpms, logprior


def tempogram_stats(env: torch.Tensor, sr: int, hop: int, start_bpm: float = 120.0):
    """Device tempogram of an onset envelope -> (tempo per frame [n], global tempo, mean tempogram [win])."""
    lib = _lib.init(_dev_index(env))
    env = env.contiguous().float()
    n = env.numel()
    win = int(np.floor(8.0 * sr / hop))
    bpms, logprior = tempo_prior(win, sr, hop, start_bpm)
    lp = torch.from_numpy(logprior.astype(np.float32)).to(env.device)
    tg_sum = torch.empty(win, dtype=torch.float32, device=env.device)
    best = torch.empty(n, dtype=torch.int32, device=env.device)
    check(lib.ac_tempogram_stats(ptr(env), n, win, ptr(lp), ptr(tg_sum), ptr(best), stream_ptr()), "ac_tempogram_stats")
    tg_mean = (tg_sum / float(n)).cpu().numpy().astype(np.float64)
    curve = bpms[best.cpu().numpy()]
    glob = float(bpms[int(np.argmax(np.log1p(1e6 * tg_mean) + logprior))])
    return curve, glob, tg_mean


def host_beat_dp(localscore: np.ndarray, period: int, tightness: float):
    """C++ scan of the beat-tracking DP (host)."""
    lib = _lib.load()
    ls = np.ascontiguousarray(localscore, dtype=np.float32)
    n = ls.shape[0]
    backlink = np.zeros(n, dtype=np.int64)
    cumscore = np.zeros(n, dtype=np.float32)
    check(lib.ac_host_beat_dp(ls.ctypes.data, n, int(period), float(tightness), backlink.ctypes.data, cumscore.ctypes.data), "ac_host_beat_dp")
    return backlink, cumscore


# ---- batched resampler in front of the per-chunk VAD ----------------------------------
_RESAMPLE_TAPS: Dict[Tuple[int, int, str], Tuple[torch.Tensor, int, int, int, int]] = {}


def resample_filter(up: int, down: int):
    """(taps float32, n_pre_pad, n_pre_remove, up, down) exactly as scipy.signal.resample_poly designs them for float32 input:
    firwin(2*half_len+1, 1/max_rate, window=('kaiser', 5.0)).astype(float32) * up, half_len = 10*max_rate."""
    import math

    import scipy.signal

    g = math.gcd(int(up), int(down))
    up, down = int(up) // g, int(down) // g
    max_rate = max(up, down)
    half_len = 10 * max_rate
    h = scipy.signal.firwin(2 * half_len + 1, 1.0 / max_rate, window=("kaiser", 5.0)).astype(np.float32)
    h = h * np.float32(up)
    n_pre_pad = down - half_len % down
    n_pre_remove = (half_len + n_pre_pad) // down
    return h, n_pre_pad, n_pre_remove, up, down


def resample_chunks(x: torch.Tensor, chunk_lens: Sequence[int], sr_in: int = 44100, sr_out: int = 16000, *, bucket: int = 4096):
    """All chunks of ``x`` (float32 CUDA, chunks back to back) resampled sr_in -> sr_out in one launch.

    Returns (batch [n_chunks, L] zero-padded to a multiple of ``bucket`` like advanced_vad.silero_length_bucket
    (vocal_pause_detector.py:190-195), out_lens).  Semantics: scipy.signal.resample_poly(chunk, sr_out, sr_in) per chunk."""
    lib = _lib.init(_dev_index(x))
    x = x.contiguous().float()
    key = (sr_out, sr_in, str(x.device))
    if key not in _RESAMPLE_TAPS:
        h, npp, npr, up, down = resample_filter(sr_out, sr_in)
        _RESAMPLE_TAPS[key] = (torch.from_numpy(h).to(x.device), npp, npr, up, down)
    taps, npp, npr, up, down = _RESAMPLE_TAPS[key]
    n_seg = len(chunk_lens)
    offs = np.concatenate([[0], np.cumsum(np.asarray(chunk_lens, dtype=np.int64))])[:-1].astype(np.int64) if n_seg else np.zeros(0, np.int64)
    lens = np.asarray(chunk_lens, dtype=np.int64)
    if n_seg and int(offs[-1] + lens[-1]) > x.numel():
        raise ValueError("chunk lengths exceed the buffer")
    out_lens = [int(lib.ac_resample_out_len(int(l), up, down)) for l in lens]
    row = max(out_lens) if out_lens else 0
    if bucket > 0:
        row += (-row) % bucket
    row = max(row, 1)
    out = torch.empty((n_seg, row), dtype=torch.float32, device=x.device)
    if n_seg:
        check(lib.ac_resample_poly(ptr(x), offs.ctypes.data_as(C.POINTER(C.c_longlong)), lens.ctypes.data_as(C.POINTER(C.c_longlong)),
                                   n_seg, up, down, ptr(taps), taps.numel(), npp, npr, row, ptr(out), stream_ptr()), "ac_resample_poly")
    return out, out_lens


# ---- F0 / formants of the legacy pause-detector branch (pure_vocal_pause_detector.py:410-459, 961-1018)
C2_HZ = 65.40639132514966   # librosa.note_to_hz("C2")
C7_HZ = 2093.004522404789   # librosa.note_to_hz("C7")


def pyin(x: torch.Tensor, sr: int = 44100, hop_length: int = 441, fmin: float = C2_HZ, fmax: float = C7_HZ, *,
         decode: bool = True):
    """librosa.pyin(y, fmin=C2, fmax=C7, sr=sr, hop_length=hop) on the GPU.

    Returns (f0 [n_frames] with NaN where unvoiced, voiced_flag bool, voiced_prob); with ``decode=False`` only
    the YIN/candidate stage runs and f0 / voiced_flag are None."""
    lib = _lib.init(_dev_index(x))
    x = x.contiguous().float()
    n = x.numel()
    nf = int(lib.ac_pyin_frame_count(n, hop_length))
    ws = torch.empty(int(lib.ac_pyin_workspace_bytes(n, hop_length, sr, fmin, fmax)), dtype=torch.uint8, device=x.device)
    vp = torch.empty(nf, dtype=torch.float32, device=x.device)
    f0 = torch.empty(nf, dtype=torch.float32, device=x.device) if decode else None
    flag = torch.empty(nf, dtype=torch.uint8, device=x.device) if decode else None
    check(lib.ac_pyin(ptr(x), n, sr, hop_length, fmin, fmax, ptr(f0), ptr(flag), ptr(vp), ptr(ws), ws.numel(), stream_ptr()), "ac_pyin")
    return f0, (flag.bool() if decode else None), vp


def lpc_formants(x: torch.Tensor, sr: int = 44100, hop_length: int = 441, order: int = 12):
    """Dense form of _extract_formants: (mags [n_frames, 3], counts [n_frames]); frame = int(0.025 * sr)."""
    lib = _lib.init(_dev_index(x))
    x = x.contiguous().float()
    n = x.numel()
    frame = int(0.025 * sr)
    nf = int(lib.ac_lpc_frame_count(n, frame, hop_length))
    mags = torch.zeros((nf, 3), dtype=torch.float32, device=x.device)
    counts = torch.zeros(nf, dtype=torch.int32, device=x.device)
    check(lib.ac_lpc_formants(ptr(x), n, frame, hop_length, order, ptr(mags), ptr(counts), stream_ptr()), "ac_lpc_formants")
    return mags, counts


def formant_tracks(mags: np.ndarray, counts: np.ndarray):
    """The reference's three ragged lists (pure_vocal_pause_detector.py:1000-1016): a frame with k >= 1 peaks
    appends to the first k tracks only; a frame with no peak (or a failed LPC) appends 0.0 to all three."""
    mags = np.asarray(mags)
    counts = np.asarray(counts)
    tracks = []
    for j in range(3):
        keep = (counts > j) | (counts == 0)
        tracks.append(np.where(counts[keep] == 0, 0.0, mags[keep, j]).astype(mags.dtype))
    return tracks


def refine_cut_points(mix: torch.Tensor, vocal: Optional[torch.Tensor], sr: int, times: Sequence[float], *,
                      zero_cross_half: int, search: int, win: int, guard_db: float, floor_db: float,
                      use_vocal_guard_first: bool = True, enable_vocal_guard: bool = True,
                      enable_mix_guard: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """The per-point loop of finalize_cut_points (refine.py:318-371) for all points in one launch.

    ``mix`` / ``vocal``: mono float32 CUDA tensors of the same length (``vocal`` may be None).  Returns
    (guard_times, final_times) as float64 numpy arrays."""
    lib = _lib.init(_dev_index(mix))
    mix = mix.contiguous().float()
    if vocal is not None:
        vocal = vocal.contiguous().float()
        if vocal.numel() != mix.numel():
            raise _lib.AudioCutError("mix and vocal must have the same length")
    p = len(times)
    t_in = torch.tensor(np.asarray(times, dtype=np.float64), dtype=torch.float64, device=mix.device)
    out = torch.empty((2, max(p, 1)), dtype=torch.float64, device=mix.device)
    check(lib.ac_refine_cut_points(ptr(mix), ptr(vocal), mix.numel(), int(sr), ptr(t_in), p, int(zero_cross_half),
                                   int(search), int(win), float(guard_db), float(floor_db),
                                   int(bool(use_vocal_guard_first)), int(bool(enable_vocal_guard)),
                                   int(bool(enable_mix_guard)), ptr(out[0]), ptr(out[1]), stream_ptr()),
          "ac_refine_cut_points")
    host = out.cpu().numpy()
    return host[0, :p].copy(), host[1, :p].copy()


def quiet_lookup_db(wave: torch.Tensor, win: int) -> torch.Tensor:
    """QuietGuardLookup.rms_db (refine.py:161-173) of a mono float32 CUDA tensor, float64 on the device."""
    lib = _lib.init(_dev_index(wave))
    wave = wave.contiguous().float()
    out = torch.empty(wave.numel(), dtype=torch.float64, device=wave.device)
    check(lib.ac_quiet_lookup_db(ptr(wave), wave.numel(), int(win), ptr(out), stream_ptr()), "ac_quiet_lookup_db")
    return out
