"""Generate the committed golden fixtures by running the REFERENCE's own code.

Run in the dev container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

What is executed unmodified from /root/reference:
  * ``audio_cut.utils.gpu_pipeline.chunk_schedule``                     -> chunk_schedule.json
  * ``audio_cut.separation.backends.MDX23OnnxBackend.infer_chunk``       -> infer_chunk.npz
  * ``EnhancedVocalSeparator._separate_with_pipeline`` and, inside it,
    ``ChunkFeatureBuilder.add_chunk/finalize`` + ``_compute_mdd_series`` -> pipeline.npz
  * ``BPMAnalyzer.extract_bpm_features`` (``--rhythm``)                  -> rhythm.json
  * the whole cut-point chain of ``v2.2_mdd`` (``--cutchain``): separator driver, feature builder, presence markers,
    ``PureVocalPauseDetector.detect_pure_vocal_pauses``, ``SeamlessSplitter._find_no_vocal_runs`` /
    ``_finalize_and_filter_cuts_v2``, ``finalize_cut_points``            -> cutchain.json

Third-party pieces that are absent from the image are stubbed, and the stubs are
NOT the thing being pinned: ``onnxruntime`` -> a fake session computing a fixed linear
map; the external MVSEP ``Conv_TDF_net_trim_model`` -> ``oracle.mdx`` stft/istft (torch);
``librosa`` -> a shim module forwarding to ``oracle.features``.  What the fixtures pin is
the reference's own host arithmetic around those calls (padding, windowing, trimming,
cropping, stem subtraction, mono mean, halo trimming, uniform overlap average,
effective-region masks, first-wins dedupe, onset-frame union, MDD weights).
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "src"))

from oracle import features as OF  # noqa: E402
from oracle import mdx as OM  # noqa: E402
from oracle import rhythm as ORH  # noqa: E402


def install_librosa_shim(full_rhythm: bool = False):
    lib = types.ModuleType("librosa")
    feat = types.ModuleType("librosa.feature")
    rhythm = types.ModuleType("librosa.feature.rhythm")
    onset = types.ModuleType("librosa.onset")
    beat = types.ModuleType("librosa.beat")
    feat.rms = lambda y=None, frame_length=2048, hop_length=512, **k: OF.rms(y, frame_length, hop_length)[None, :]
    feat.spectral_flatness = lambda y=None, hop_length=512, n_fft=2048, **k: OF.spectral_flatness(y, n_fft, hop_length)[None, :]
    onset.onset_strength = lambda y=None, sr=22050, hop_length=512, aggregate=np.mean, **k: OF.onset_strength(
        y, sr, hop_length, aggregate=aggregate
    )
    onset.onset_detect = lambda onset_envelope=None, sr=22050, hop_length=512, **k: OF.onset_detect(onset_envelope, sr, hop_length)
    lib.frames_to_time = lambda frames, sr=22050, hop_length=512, **k: np.asarray(frames) * hop_length / float(sr)
    if not full_rhythm:  # the round-1 fixtures (pipeline.npz) were generated with these placeholders
        rhythm.tempo = lambda onset_envelope=None, **k: np.full(len(onset_envelope), 120.0)

        def beat_track(y=None, onset_envelope=None, **k):
            if y is not None:
                raise RuntimeError("shim: beat_track(y=...) not provided")  # BPMAnalyzer falls back to defaults
            return 120.0, np.zeros(0, dtype=int)
    else:
        def _tempo(onset_envelope=None, sr=22050, hop_length=512, aggregate=np.mean, start_bpm=120.0, **k):
            return ORH.tempo(onset_envelope, sr, hop_length, start_bpm=start_bpm, aggregate=None if aggregate is None else "mean")

        rhythm.tempo = _tempo

        def beat_track(y=None, onset_envelope=None, sr=22050, hop_length=512, start_bpm=120.0, tightness=100, **k):
            return ORH.beat_track(onset_envelope, sr, hop_length, y=y, start_bpm=start_bpm, tightness=float(tightness))

    beat.beat_track = beat_track
    feat.rhythm = rhythm
    lib.feature, lib.onset, lib.beat = feat, onset, beat
    lib.load = None
    lib.resample = None
    for m in (lib, feat, rhythm, onset, beat):
        sys.modules[m.__name__] = m
    # the orchestrator module imports its I/O libraries at module level; nothing here calls them
    for name in ("soundfile", "pydub"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["pydub"], "AudioSegment"):
        sys.modules["pydub"].AudioSegment = object


def gen_chunk_schedule():
    from audio_cut.utils.gpu_pipeline import chunk_schedule

    cases = []
    args = [
        (30.0, 10.0, 2.5, 0.5), (240.0, 10.0, 2.5, 0.5), (3600.0, 10.0, 2.5, 0.5), (9.99, 10.0, 2.5, 0.5),
        (10.0, 10.0, 2.5, 0.5), (10.0000005, 10.0, 2.5, 0.5), (10.001, 10.0, 2.5, 0.5), (0.0, 10.0, 2.5, 0.5),
        (47.3, 10.0, 2.5, 0.5), (61.234567, 7.0, 6.9, 4.0), (100.0, 5.0, 0.0, 0.0), (33.3, 0.05, 0.01, 1.0),
        (17.5, 10.0, 2.5, 0.5), (25.0, 10.0, 2.5, 0.5), (1234.5678, 12.5, 3.3, 0.7), (23.7, 10.0, 2.5, 0.5),
    ]
    for total, c, o, h in args:
        plans = chunk_schedule(total, chunk_s=c, overlap_s=o, halo_s=h)
        cases.append(
            {
                "args": [total, c, o, h],
                "plans": [[p.index, repr(p.start_s), repr(p.end_s), repr(p.halo_left_s), repr(p.halo_right_s)] for p in plans],
            }
        )
    with open(os.path.join(HERE, "chunk_schedule.json"), "w") as f:
        json.dump(cases, f, indent=0)
    print("chunk_schedule.json", len(cases), "cases")


SMALL = OM.MdxGeometry(n_fft=512, hop=128, dim_f=224, dim_t=32)  # chunk_size 3968, gen 3456
CH_GAIN = np.array([0.5, 0.4, 0.3, 0.45], np.float32)


class _FakeChunkModel:
    n_fft = SMALL.n_fft
    chunk_size = SMALL.chunk_size

    def stft(self, x):
        return OM.stft(x, SMALL)

    def istft(self, y):
        return OM.istft(y, SMALL)

    def eval(self):
        return self


class _FakeSession:
    def run(self, _names, feeds):
        (x,) = feeds.values()
        return [(x * CH_GAIN[None, :, None, None]).astype(np.float32)]


def gen_infer_chunk():
    from audio_cut.separation.backends import MDX23OnnxBackend

    out = {}
    rng = np.random.default_rng(7)
    for tag, shape, otype in (("mono", (10007,), "vocal"), ("stereo", (2, 7000), "instrumental"), ("short", (300,), "vocal")):
        be = MDX23OnnxBackend("/nonexistent", provider="CPUExecutionProvider", execution_device="cpu", align_hop=256)
        be._session, be._chunk_model = _FakeSession(), _FakeChunkModel()
        be._onnx_input, be._onnx_output = "input", "output"
        be._resolved_output_type = otype
        x = (0.3 * rng.standard_normal(shape)).astype(np.float32)
        res = be.infer_chunk(x)
        out[f"{tag}_in"], out[f"{tag}_vocal"], out[f"{tag}_instr"] = x, res.vocal, res.instrumental
        out[f"{tag}_otype"] = np.array(otype)
    np.savez_compressed(os.path.join(HERE, "infer_chunk.npz"), **out)
    print("infer_chunk.npz", {k: getattr(v, "shape", None) for k, v in out.items()})


def gen_pipeline():
    install_librosa_shim()
    from audio_cut.separation.backends import IVocalSeparatorBackend, SeparationOutputs
    from audio_cut.utils.gpu_pipeline import PipelineConfig
    from vocal_smart_splitter.core import enhanced_vocal_separator as evs

    sr = 2000

    class FakeBackend(IVocalSeparatorBackend):
        calls = 0

        def load_model(self):
            pass

        def sample_rate(self):
            return sr

        def infer_chunk(self, mix_chunk, **kw):
            k = FakeBackend.calls
            FakeBackend.calls += 1
            n = mix_chunk.shape[-1]
            ramp = np.linspace(0.0, 1.0, n, dtype=np.float32)
            v = (0.7 * mix_chunk + np.float32(0.01 * (k + 1)) * ramp).astype(np.float32)
            return SeparationOutputs(vocal=v, instrumental=(mix_chunk - v).astype(np.float32))

    sep = object.__new__(evs.EnhancedVocalSeparator)
    sep.sample_rate = sr
    sep._pipeline_cfg = PipelineConfig(enable=False)
    rng = np.random.default_rng(11)
    n = int(23.7 * sr)
    t = np.arange(n) / sr
    audio = (0.2 * np.sin(2 * np.pi * 110 * t) * (np.sin(2 * np.pi * 0.4 * t) > 0) + 0.05 * rng.standard_normal(n)).astype(np.float32)
    ctx = sep._build_cpu_context(n / float(sr))
    vocal, instr, cache, vad = sep._separate_with_pipeline(audio, FakeBackend(), ctx)
    np.savez_compressed(
        os.path.join(HERE, "pipeline.npz"),
        sr=sr, audio=audio, vocal=vocal, instr=instr,
        rms_series=cache.rms_series, spectral_flatness=cache.spectral_flatness, onset_envelope=cache.onset_envelope,
        onset_frames=cache.onset_frames, mdd_series=cache.mdd_series, global_mdd=cache.global_mdd,
        rms_max=cache.rms_max, onset_max=cache.onset_max, hop_length=cache.hop_length, duration_s=cache.duration_s,
        n_chunks=len(ctx.plans), processed=ctx.gpu_meta["gpu_pipeline_processed_chunks"],
    )
    print("pipeline.npz", vocal.shape, cache.rms_series.shape, cache.onset_frames, ctx.gpu_meta)


def gen_rhythm():
    """BPMAnalyzer.extract_bpm_features of the REFERENCE (adaptive_vad_enhancer.py:48-298, unmodified) on top of the shim
    whose librosa.beat / librosa.feature.rhythm forward to oracle.rhythm: pins the classification, stability, variance
    and adaptive-factor arithmetic (the librosa algorithms themselves stay 'parity unpinned', see oracle/rhythm.py)."""
    install_librosa_shim(full_rhythm=True)
    from vocal_smart_splitter.core.adaptive_vad_enhancer import BPMAnalyzer

    sys.path.insert(0, ROOT)
    from audio_cut_b200 import synth

    cases = []
    sr = 44100
    for tag, audio in (("track20", synth.synth_track(20.0, seed=2, stereo=False)), ("song24", synth.synth_song(24.0, seed=1)),
                       ("clicks150", _click_wave(16.0, 150.0, sr)), ("clicks72", _click_wave(20.0, 72.0, sr)), ("silence", np.zeros(sr * 6, np.float32))):
        bf = BPMAnalyzer(sr).extract_bpm_features(audio)
        cases.append({"tag": tag, "main_bpm": float(np.squeeze(bf.main_bpm)), "bpm_category": bf.bpm_category,
                      "beat_strength": float(bf.beat_strength), "bpm_confidence": float(bf.bpm_confidence),
                      "tempo_variance": float(bf.tempo_variance),
                      "adaptive_factors": {k: (float(v) if isinstance(v, (int, float, np.floating)) and not isinstance(v, bool) else v)
                                           for k, v in (bf.adaptive_factors or {}).items()},
                      "beat_positions": [int(b) for b in np.asarray(bf.beat_positions).ravel()]})
    with open(os.path.join(HERE, "rhythm.json"), "w") as f:
        json.dump(cases, f, indent=0)
    print("rhythm.json", [(c["tag"], round(c["main_bpm"], 2), c["bpm_category"], len(c["beat_positions"])) for c in cases])


def _click_wave(seconds, bpm, sr):
    n = int(seconds * sr)
    x = np.zeros(n, np.float32)
    bl = int(0.02 * sr)
    burst = (np.sin(2 * np.pi * 1000.0 * np.arange(bl) / sr) * np.exp(-np.arange(bl) / (0.004 * sr))).astype(np.float32)
    step = 60.0 / bpm * sr
    k = 0
    while int(k * step) + bl < n:
        p = int(k * step)
        x[p:p + bl] += burst
        k += 1
    return x + 1e-4 * np.random.default_rng(5).standard_normal(n).astype(np.float32)


CUTCHAIN_SECONDS = 40.0
CUTCHAIN_N_FFT = 6144  # the reference hard-codes it (backends.py:264)


def gen_cutchain(small: bool = False):
    """SURVEY.md X1 fixture (``small``: the backend is a fixed linear map instead of the network, so that the CPU-only test
    suite can rebuild the stems in no time - what that variant pins is the restated detector chain, not the separation): the reference's OWN chain - EnhancedVocalSeparator._separate_with_pipeline (oracle network as the
    backend, Kim_Vocal geometry) -> ChunkFeatureBuilder -> VocalSeparator._compute_vocal_presence_markers ->
    PureVocalPauseDetector.detect_pure_vocal_pauses -> candidate assembly of _process_pure_vocal_split ->
    SeamlessSplitter._finalize_and_filter_cuts_v2 / finalize_cut_points - all unmodified, on the oracle-backed librosa shim.
    Every configuration value the chain reads is recorded.  The stems themselves are not stored (tests recompute them with
    the oracle from the seeded input); their checksums are."""
    install_librosa_shim(full_rhythm=True)
    import logging

    logging.disable(logging.CRITICAL)
    from audio_cut.separation.backends import IVocalSeparatorBackend, SeparationOutputs
    from audio_cut.utils.gpu_pipeline import PipelineConfig
    from vocal_smart_splitter.core import enhanced_vocal_separator as evs
    from vocal_smart_splitter.core import pure_vocal_pause_detector as pv
    from vocal_smart_splitter.core import seamless_splitter as ss
    from vocal_smart_splitter.core.vocal_separator import VocalSeparator
    from vocal_smart_splitter.utils import config_manager as cm

    from audio_cut_b200 import synth, unet_weights as uw
    from oracle import unet as ounet

    recorded = {}
    real_get = cm.get_config

    def spy(key, default=None):
        val = real_get(key, default)
        recorded[key] = val
        return val

    for mod in (cm, pv, ss, evs):
        if hasattr(mod, "get_config"):
            mod.get_config = spy
    import vocal_smart_splitter.core.vocal_separator as vsm
    import vocal_smart_splitter.core.vocal_pause_detector as vpd

    for mod in (vsm, vpd):
        if hasattr(mod, "get_config"):
            mod.get_config = spy

    sr = 44100
    seconds = 48.0 if small else CUTCHAIN_SECONDS
    audio = synth.synth_song(seconds, sr=sr, seed=3 if small else 0)
    geo = uw.UNetGeometry()
    net = None if small else ounet.build_net(uw.random_state(geo), geo.dim_f, geo.dim_t, geo.g)
    mg = OM.MdxGeometry(CUTCHAIN_N_FFT, 1024, 3072, 256)
    align_hop = 4096

    class OracleBackend(IVocalSeparatorBackend):
        def load_model(self):
            pass

        def sample_rate(self):
            return sr

        def infer_chunk(self, mix_chunk, **kw):
            if small:  # tests/helpers.py:linear_backend is the same map
                v = (np.float32(0.7) * mix_chunk).astype(np.float32)
                return SeparationOutputs(vocal=v, instrumental=(mix_chunk - v).astype(np.float32))
            v, i = OM.infer_chunk(mix_chunk, net, mg, align_hop=align_hop)
            return SeparationOutputs(vocal=v, instrumental=i)

    sep = object.__new__(evs.EnhancedVocalSeparator)
    sep.sample_rate = sr
    sep._pipeline_cfg = PipelineConfig(enable=False)
    ctx = sep._build_cpu_context(len(audio) / float(sr))
    vocal, instr, cache, vad = sep._separate_with_pipeline(audio, OracleBackend(), ctx)
    markers = VocalSeparator(sr)._compute_vocal_presence_markers(vocal)
    det = pv.PureVocalPauseDetector(sr)
    pauses = det.detect_pure_vocal_pauses(vocal, enable_mdd_enhancement=True, original_audio=audio, feature_cache=cache, vad_segments=vad)
    assert not det._last_focus_windows, "the restated chain assumes no Silero model (empty focus windows)"
    cands = [(float(p.cut_point), float(p.confidence)) for p in pauses]
    fake = types.SimpleNamespace(sample_rate=sr, _set_guard_adjustments=lambda a: None)
    min_pure = float(spy("quality_control.pure_music_min_duration", 0.0))
    spans = ss.SeamlessSplitter._find_no_vocal_runs(fake, vocal, min_pure) if min_pure > 0 else []
    for a, b in spans:
        cands += [(float(a), 1.0), (float(b), 1.0)]
    marker_times = [float(t) for t in markers.get("vocal_presence_cut_points_sec", [])]
    duration = len(audio) / sr
    protected = set()
    for t in marker_times:
        if t <= 0.0 or t >= duration:
            continue
        cands.append((t, 1.0))
        protected.add(int(round(t * sr)))
    res = ss.SeamlessSplitter._finalize_and_filter_cuts_v2(fake, list(cands), audio, pure_vocal_audio=vocal)
    final = set(int(b) for b in res.sample_boundaries)
    for s_ in protected:
        s_ = int(min(max(s_, 0), len(audio)))
        if s_ not in (0, len(audio)):
            final.add(s_)

    def js(v):
        if isinstance(v, dict):
            return {k: js(x) for k, x in v.items()}
        if isinstance(v, (list, tuple)):
            return [js(x) for x in v]
        if isinstance(v, (np.floating, np.integer)):
            return v.item()
        return v

    bf = cache.bpm_features
    out = {
        "seconds": seconds, "n_fft": mg.n_fft, "hop": mg.hop, "dim_f": mg.dim_f, "dim_t": mg.dim_t, "g": geo.g, "align_hop": align_hop,
        "song_seed": 3 if small else 0, "sr": sr, "n_chunks": len(ctx.plans),
        "config": js(recorded),
        "vocal_sum": repr(float(np.sum(vocal.astype(np.float64)))), "vocal_abs_sum": repr(float(np.sum(np.abs(vocal.astype(np.float64))))),
        "cache": {"global_mdd": repr(float(cache.global_mdd)), "main_bpm": repr(float(np.squeeze(bf.main_bpm))), "bpm_category": bf.bpm_category,
                  "rms_max": repr(float(cache.rms_max)), "onset_max": repr(float(cache.onset_max)), "n_frames": int(cache.frame_count()),
                  "onset_frames": [int(x) for x in cache.onset_frames], "beat_times": [repr(float(x)) for x in np.asarray(cache.beat_times)]},
        "marker_times": [repr(t) for t in marker_times],
        "pauses": [[repr(float(p.start_time)), repr(float(p.end_time)), repr(float(p.cut_point)), repr(float(p.confidence)), p.quality_grade, p.pause_type]
                   for p in pauses],
        "pure_music_spans": [[repr(float(a)), repr(float(b))] for a, b in spans],
        "candidates": [[repr(t), repr(s)] for t, s in cands],
        "refined_boundaries": [int(b) for b in res.sample_boundaries],
        "final_times": [repr(float(p.t)) for p in res.final_points],
        "adjustments": [[repr(float(a.raw_time)), repr(float(a.guard_time)), repr(float(a.final_time))] for a in res.adjustments],
        "sample_boundaries": sorted(final),
    }
    name = "cutchain_small.json" if small else "cutchain.json"
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(out, f, indent=0)
    print(name, ": pauses", len(pauses), "candidates", len(cands), "spans", len(spans), "markers", marker_times,
          "boundaries", out["sample_boundaries"], "config keys", len(recorded))


def main():
    torch.set_num_threads(os.cpu_count() or 4)
    if "--rhythm" in sys.argv:
        return gen_rhythm()
    if "--cutchain" in sys.argv:
        return gen_cutchain()
    if "--cutchain-small" in sys.argv:
        return gen_cutchain(small=True)
    if "--cuts44k" in sys.argv:
        return gen_cuts_44k()
    if "--cuts" not in sys.argv:
        gen_chunk_schedule()
        gen_infer_chunk()
        gen_pipeline()
    gen_cuts()
    gen_cuts_44k()


def _load_ref_refine():
    import importlib.util

    spec = importlib.util.spec_from_file_location("ref_refine", os.path.join(REF, "src/audio_cut/cutting/refine.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_refine"] = mod
    spec.loader.exec_module(mod)
    return mod


def gen_cuts_44k():
    """finalize_cut_points of the reference at 44.1 / 22.05 kHz with its default guard geometry, including every
    CutAdjustment (raw / guard / final time), stereo mixes, a missing vocal stem, digital silence, points at the
    track ends and disabled guards: the fixture of the GPU refinement (audio_cut_b200.refine)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import cut_case

    mod = _load_ref_refine()
    cases = []
    for seed in range(8):
        mix, vocal, sr, pts, kw = cut_case(seed)
        res = mod.finalize_cut_points(mod.CutContext(sr=sr, mix_wave=mix, vocal_wave=vocal),
                                      [mod.CutPoint(t=a, score=b) for a, b in pts], **kw)
        cases.append({"seed": seed, "sample_boundaries": [int(v) for v in res.sample_boundaries],
                      "final_times": [repr(float(p.t)) for p in res.final_points],
                      "adjustments": [[repr(float(a.raw_time)), repr(float(a.guard_time)), repr(float(a.final_time))]
                                      for a in res.adjustments],
                      "n_suppressed": len(res.suppressed_points)})
    with open(os.path.join(HERE, "cuts_44k.json"), "w") as f:
        json.dump(cases, f, indent=0)
    print("cuts_44k.json", [(len(c["sample_boundaries"]), sum(a[0] != a[2] for a in c["adjustments"])) for c in cases])


def gen_cuts():
    """finalize_cut_points of the reference (refine.py, loaded standalone: its package __init__ pulls librosa)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("ref_refine", os.path.join(REF, "src/audio_cut/cutting/refine.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_refine"] = mod
    spec.loader.exec_module(mod)
    cases = []
    for seed in range(4):
        rng = np.random.default_rng(100 + seed)
        sr = 8000
        n = int(24.0 * sr)
        t = np.arange(n) / sr
        gate = (np.sin(2 * np.pi * 0.23 * t + seed) > -0.3).astype(np.float64)
        vocal = (0.3 * np.sin(2 * np.pi * 180 * t) * gate + 1e-4 * rng.standard_normal(n)).astype(np.float32)
        mix = (vocal + 0.1 * np.sin(2 * np.pi * 55 * t) * (np.sin(2 * np.pi * 0.11 * t) > 0) + 1e-3 * rng.standard_normal(n)).astype(np.float32)
        pts = [(float(x), float(s)) for x, s in zip(rng.uniform(0.2, 23.8, 14), rng.uniform(0, 1, 14))]
        kw = dict(min_gap_s=1.0, guard_db=1.5, search_right_ms=450.0, guard_win_ms=80.0, floor_db=-45.0 + 5 * seed,
                  zero_cross_win_ms=8.0, min_boundary_s=0.5, topk_per_10s=(None if seed % 2 else 4))
        res = mod.finalize_cut_points(mod.CutContext(sr=sr, mix_wave=mix, vocal_wave=vocal), [mod.CutPoint(t=a, score=b) for a, b in pts], **kw)
        cases.append({"seed": 100 + seed, "points": pts, "kwargs": kw, "sample_boundaries": [int(v) for v in res.sample_boundaries],
                      "final_times": [repr(float(p.t)) for p in res.final_points]})
    with open(os.path.join(HERE, "cuts.json"), "w") as f:
        json.dump(cases, f, indent=0)
    print("cuts.json", [c["sample_boundaries"] for c in cases])


if __name__ == "__main__":
    main()
