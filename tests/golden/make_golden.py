"""Generate the committed golden fixtures by running the REFERENCE's own code.

Run in the dev container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

What is executed unmodified from /root/reference:
  * ``audio_cut.utils.gpu_pipeline.chunk_schedule``                     -> chunk_schedule.json
  * ``audio_cut.separation.backends.MDX23OnnxBackend.infer_chunk``       -> infer_chunk.npz
  * ``EnhancedVocalSeparator._separate_with_pipeline`` and, inside it,
    ``ChunkFeatureBuilder.add_chunk/finalize`` + ``_compute_mdd_series`` -> pipeline.npz

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


def install_librosa_shim():
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
    rhythm.tempo = lambda onset_envelope=None, **k: np.full(len(onset_envelope), 120.0)

    def beat_track(y=None, onset_envelope=None, **k):
        if y is not None:
            raise RuntimeError("shim: beat_track(y=...) not provided")  # BPMAnalyzer falls back to defaults
        return 120.0, np.zeros(0, dtype=int)

    beat.beat_track = beat_track
    feat.rhythm = rhythm
    lib.feature, lib.onset, lib.beat = feat, onset, beat
    lib.load = None
    for m in (lib, feat, rhythm, onset, beat):
        sys.modules[m.__name__] = m


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


def main():
    torch.set_num_threads(4)
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
