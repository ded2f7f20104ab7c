"""The reference-facing drop-ins (separator / backend / feature builder) end to end on the GPU."""
import numpy as np
import pytest
import torch

from helpers import sdr_db

pytestmark = pytest.mark.gpu


def _small_backend(precision="fp32", stereo_sr=8000):
    from audio_cut_b200 import unet_weights as uw
    from audio_cut_b200.backends import B200Mdx23Backend

    geo = uw.UNetGeometry(dim_f=256, dim_t=32, g=16)
    st = uw.random_state(geo)
    be = B200Mdx23Backend(weights=st, geometry=geo, n_fft=640, hop=128, align_hop=256, precision=precision, output_type="vocal")
    be.load_model()
    return be, st, geo


def test_backend_infer_chunk_matches_oracle():
    from audio_cut_b200 import synth
    from oracle import mdx
    from oracle import unet as ounet

    be, st, geo = _small_backend()
    ref_net = ounet.build_net(st, geo.dim_f, geo.dim_t, geo.g)
    mg = mdx.MdxGeometry(640, 128, 256, 32)
    audio = synth.synth_track(2.5, sr=8000, seed=2)
    for chunk in (audio, audio[0], audio[0][:777]):
        out = be.infer_chunk(chunk)
        rv, ri = mdx.infer_chunk(chunk, ref_net, mg, align_hop=256, output_type="vocal")
        assert out.vocal.shape == rv.shape and out.vocal.dtype == np.float32
        assert sdr_db(rv, out.vocal) > 60 and sdr_db(ri, out.instrumental) > 60
    m = be.get_performance_metrics(reset=True)
    assert m["chunks"] == 3 and m["compute_ms"] > 0
    assert be.describe_input() == {"name": "input", "shape": [1, 4, 256, 32]}
    with pytest.raises(Exception):
        be.fallback_to_cpu()


def test_separator_dropin_matches_oracle_pipeline():
    from audio_cut_b200 import synth
    from audio_cut_b200.gpu_pipeline import PipelineConfig
    from audio_cut_b200.separator import B200VocalSeparator, SeparationResult
    from oracle import mdx, pipeline, planner
    from oracle import unet as ounet

    sr = 8000
    be, st, geo = _small_backend()
    cfg = PipelineConfig(chunk_s=4.0, overlap_s=1.0, halo_s=0.25, align_hop=256)
    sep = B200VocalSeparator(sr, backend=be, pipeline_config=cfg)
    audio = synth.synth_track(17.3, sr=sr, seed=4, stereo=False)  # the pipeline is mono-in
    res = sep.separate_for_detection(audio)
    assert isinstance(res, SeparationResult) and res.backend_used == "B200Mdx23Backend"
    assert res.vocal_track.shape == audio.shape and res.vocal_track.dtype == np.float32

    ref_net = ounet.build_net(st, geo.dim_f, geo.dim_t, geo.g)
    mg = mdx.MdxGeometry(640, 128, 256, 32)
    plans = planner.chunk_schedule(len(audio) / float(sr), 4.0, 1.0, 0.25)
    rv, ri = pipeline.separate_track(audio, lambda ch: mdx.infer_chunk(ch, ref_net, mg, align_hop=256), sr=sr, plans=plans)
    assert sdr_db(rv, res.vocal_track) > 60 and sdr_db(ri, res.instrumental_track) > 60

    # feature cache vs the oracle's restatement of ChunkFeatureBuilder
    cf = pipeline.ChunkFeatures(sr)
    for p in plans:
        cs, ce, _, _ = planner.sample_bounds(p, sr, len(audio))
        cf.add_chunk(p, audio[cs:ce])
    ref = cf.finalize()
    fc = res.feature_cache
    assert fc.hop_length == cf.hop_length and fc.rms_series.shape == ref["rms_series"].shape
    np.testing.assert_allclose(fc.rms_series, ref["rms_series"], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(fc.spectral_flatness, ref["spectral_flatness"], rtol=1e-4, atol=1e-9)
    np.testing.assert_allclose(fc.onset_envelope, ref["onset_envelope"], rtol=1e-4, atol=1e-4 * ref["onset_envelope"].max())
    np.testing.assert_array_equal(fc.onset_frames, ref["onset_frames"])
    np.testing.assert_allclose(fc.mdd_series, ref["mdd_series"], rtol=2e-4, atol=1e-6)
    assert abs(fc.global_mdd - ref["global_mdd"]) < 1e-4
    assert fc.bpm_features is not None and fc.tempo_curve.shape == fc.rms_series.shape

    meta = res.gpu_meta
    for k in ("gpu_pipeline_enabled", "gpu_pipeline_used", "gpu_pipeline_device", "gpu_pipeline_chunks",
              "gpu_pipeline_processed_chunks", "gpu_pipeline_h2d_ms", "gpu_pipeline_dtoh_ms", "gpu_pipeline_compute_ms",
              "gpu_pipeline_peak_mem_bytes", "gpu_pipeline_chunk_invocations", "mdx23_output_type", "silero_vad_segments",
              "gpu_pipeline_mdx23_input", "gpu_pipeline_device_name", "gpu_pipeline_config"):
        assert k in meta, k
    assert meta["gpu_pipeline_used"] is True and meta["gpu_pipeline_processed_chunks"] == len(plans)
    assert set(res.quality_metrics) == {"vocal_presence_cut_points_sec", "vocal_presence_cut_points_samples",
                                        "vocal_presence_segments", "pure_music_segments"}


def test_cut_refinement_on_device_resident_stems():
    """separate_for_detection -> finalize_cut_points with the stems still in HBM: the cut samples equal the oracle's
    finalize_cut_points run on the host copies of the same stems (bit-exact sample boundaries)."""
    from audio_cut_b200 import refine as R, synth
    from audio_cut_b200.gpu_pipeline import PipelineConfig
    from audio_cut_b200.separator import B200VocalSeparator
    from oracle import cuts

    sr = 8000
    be, _, _ = _small_backend()
    sep = B200VocalSeparator(sr, backend=be, pipeline_config=PipelineConfig(chunk_s=4.0, overlap_s=1.0, halo_s=0.25, align_hop=256))
    audio = synth.synth_track(17.3, sr=sr, seed=9, stereo=False)
    res = sep.separate_for_detection(audio)
    dev = sep.last_device_stems()
    np.testing.assert_array_equal(dev["vocal"].cpu().numpy(), res.vocal_track)
    np.testing.assert_array_equal(dev["mix"].cpu().numpy(), audio)
    rng = np.random.default_rng(2)
    pts = [(float(a), float(b)) for a, b in zip(rng.uniform(0.2, 17.0, 24), rng.uniform(0, 1, 24))]
    kw = dict(min_gap_s=0.5, floor_db=-30.0, guard_db=1.0)
    got = R.finalize_cut_points(R.CutContext(sr=sr, mix_wave=dev["mix"], vocal_wave=dev["vocal"]),
                                [R.CutPoint(a, b) for a, b in pts], **kw)
    bounds, times = cuts.finalize_cut_points(audio, res.vocal_track, sr, pts, **kw)
    assert got.sample_boundaries == bounds and [p.t for p in got.final_points] == times and len(times) >= 8


def test_pinned_input_and_output_arrays():
    """A page-locked caller array goes to the device without the staging copy; results equal the pageable path, are
    float32 numpy arrays of the input length that stay valid (and writable) after later calls."""
    from audio_cut_b200 import _lib, synth
    from audio_cut_b200.gpu_pipeline import PipelineConfig
    from audio_cut_b200.separator import B200VocalSeparator

    sr = 8000
    be, _, _ = _small_backend()
    sep = B200VocalSeparator(sr, backend=be, pipeline_config=PipelineConfig(chunk_s=4.0, overlap_s=1.0, halo_s=0.25, align_hop=256))
    audio = synth.synth_track(9.7, sr=sr, seed=6, stereo=False)
    pinned = torch.empty(audio.shape, dtype=torch.float32, pin_memory=True).numpy()
    pinned[...] = audio
    lib = _lib.load()
    assert lib.ac_host_is_pinned(pinned.ctypes.data, pinned.nbytes) == 1
    assert lib.ac_host_is_pinned(audio.ctypes.data, audio.nbytes) == 0
    a = sep.separate_for_detection(audio)
    va, ia = a.vocal_track.copy(), a.instrumental_track.copy()
    b = sep.separate_for_detection(pinned)
    np.testing.assert_array_equal(b.vocal_track, va)
    np.testing.assert_array_equal(b.instrumental_track, ia)
    for _ in range(3):  # later calls must not recycle memory the caller still holds
        sep.separate_for_detection(audio * np.float32(0.5))
    np.testing.assert_array_equal(a.vocal_track, va)
    np.testing.assert_array_equal(b.instrumental_track, ia)
    assert a.vocal_track.dtype == np.float32 and a.vocal_track.shape == audio.shape and a.vocal_track.flags.writeable
    sep.pinned_outputs = False
    c = sep.separate_for_detection(audio)
    np.testing.assert_array_equal(c.vocal_track, va)


def test_vad_hook_is_called_per_chunk():
    from audio_cut_b200 import synth
    from audio_cut_b200.gpu_pipeline import PipelineConfig
    from audio_cut_b200.separator import B200VocalSeparator

    be, st, geo = _small_backend()
    from audio_cut_b200.chunk_vad import B200ChunkVAD
    from oracle import mdx
    from oracle import unet as ounet

    ref_net = ounet.build_net(st, geo.dim_f, geo.dim_t, geo.g)
    mg = mdx.MdxGeometry(640, 128, 256, 32)

    seen, chunks = [], []

    def vad(plan, vocal_chunk, sr):
        seen.append((plan.index, len(vocal_chunk)))
        chunks.append(np.array(vocal_chunk))
        return [{"start": plan.start_s, "end": plan.end_s}]

    made = []

    def factory(sr):
        made.append(B200ChunkVAD(sr, inference_fn=lambda x: [{"start": 0, "end": len(x)}]))
        return made[-1]

    cfg = PipelineConfig(chunk_s=4.0, overlap_s=1.0, halo_s=0.25, align_hop=256)
    sep = B200VocalSeparator(8000, backend=be, pipeline_config=cfg, vad_fn=vad, chunk_vad=factory)
    audio = synth.synth_track(9.0, sr=8000, stereo=False)
    res = sep.separate_for_detection(audio)
    assert [i for i, _ in seen] == list(range(len(seen)))
    assert res.gpu_meta["silero_vad_segments"] == len(res.vad_segments)
    # the hook sees each chunk's OWN output (whole chunk, halos included, before the overlap average):
    # enhanced_vocal_separator.py:412-417 passes outputs.vocal of infer_chunk
    from audio_cut_b200.gpu_pipeline import chunk_schedule

    plans = chunk_schedule(len(audio) / 8000.0, chunk_s=4.0, overlap_s=1.0, halo_s=0.25)
    assert len(plans) == len(chunks) >= 2
    for p, got in zip(plans, chunks):
        cs, ce, _, _ = p.sample_bounds(8000, len(audio))
        assert len(got) == ce - cs
        ref_v, _ = mdx.infer_chunk(audio[cs:ce], ref_net, mg, align_hop=256)
        assert sdr_db(ref_v, got) > 60
    # interior chunks differ from the stitched track inside the seams (the stitched samples are averages)
    cs, ce, es, ee = plans[1].sample_bounds(8000, len(audio))
    assert not np.allclose(chunks[1][: es - cs + 100], res.vocal_track[cs : es + 100], atol=1e-7)
    # the SileroChunkVAD-style adapter: every chunk reports all-speech -> one merged span over the whole track
    merged = made[0].finalize()
    assert len(merged) == 1 and abs(merged[0]["start"]) < 1e-9 and abs(merged[0]["end"] - len(audio) / 8000.0) < 1e-3
    assert len(res.vad_segments) == len(seen) + 1


def test_vocal_features_dropin_matches_oracle():
    """PureVocalPauseDetector._extract_vocal_features (pure_vocal_pause_detector.py:410-459) on the GPU."""
    from audio_cut_b200 import synth
    from audio_cut_b200.vocal_features import extract_vocal_features
    from oracle import features as OF

    y = synth.synth_track(3.0, seed=11)[0].astype(np.float32)
    vf = extract_vocal_features(y, 44100, 441)
    n_frames = 1 + len(y) // 441
    for a in (vf.f0_contour, vf.f0_confidence, vf.spectral_centroid, vf.harmonic_ratio, vf.zero_crossing_rate, vf.rms_energy):
        assert a.shape == (n_frames,)
    assert len(vf.formant_energies) == 3
    np.testing.assert_allclose(vf.rms_energy, OF.rms(y, 2048, 441), rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(vf.zero_crossing_rate, OF.zero_crossing_rate(y, 2048, 441), atol=1e-7)
    np.testing.assert_allclose(vf.spectral_centroid, OF.spectral_centroid(y, 44100, 2048, 441), rtol=1e-4, atol=1e-2)
    np.testing.assert_allclose(vf.harmonic_ratio, OF.low_band_ratio(y, 2048, 441), rtol=1e-4, atol=1e-7)
    r_f0, r_flag, r_vp = OF.pyin(y, 44100, hop_length=441)
    assert np.mean(np.isnan(vf.f0_contour) == np.isnan(r_f0)) > 0.99
    assert np.mean(np.abs(vf.f0_confidence - r_vp) < 1e-3) > 0.95
