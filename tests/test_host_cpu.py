"""CPU-only checks: host logic of the drop-in, C-ABI surface, oracle cross-checks."""
import json
import os
import re

import numpy as np
import pytest

from oracle import features as OF
from oracle import planner

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_chunk_schedule_host_mirror_matches_reference_golden(golden_dir):
    from audio_cut_b200.gpu_pipeline import chunk_schedule

    for case in json.load(open(os.path.join(golden_dir, "chunk_schedule.json"))):
        t, c, o, h = case["args"]
        plans = chunk_schedule(t, chunk_s=c, overlap_s=o, halo_s=h)
        got = [[p.index, repr(p.start_s), repr(p.end_s), repr(p.halo_left_s), repr(p.halo_right_s)] for p in plans]
        assert got == case["plans"]


def test_sample_bounds_match_oracle():
    from audio_cut_b200.gpu_pipeline import chunk_schedule

    for total in (1_323_000, 10_584_000, 441_001, 330_750):
        a = chunk_schedule(total / 44100.0)
        b = planner.chunk_schedule(total / 44100.0)
        assert [p.sample_bounds(44100, total) for p in a] == [planner.sample_bounds(p, 44100, total) for p in b]


def test_pipeline_config_mapping_and_meta_keys():
    from audio_cut_b200.gpu_pipeline import (InflightLimiter, PinnedBufferPool, PipelineConfig, PipelineContext, Streams,
                                             chunk_schedule)

    cfg = PipelineConfig.from_mapping({"chunk_seconds": 8, "overlap_seconds": 2, "halo_seconds": 0.25, "align_hop": 2048,
                                       "strict_mode": True, "ort": {"graph_optimization_level": "basic"}})
    assert (cfg.chunk_s, cfg.overlap_s, cfg.halo_s, cfg.align_hop) == (8.0, 2.0, 0.25, 2048)
    ctx = PipelineContext(device="cuda:0", streams=Streams(), plans=chunk_schedule(30.0), pinned_pool=None,
                          limiter=InflightLimiter(2), config=cfg, use_streams=True)
    ctx.mark_failure("separation", "boom")
    meta = ctx.to_meta()
    for k in ("gpu_pipeline_enabled", "gpu_pipeline_used", "gpu_pipeline_device", "gpu_pipeline_chunks", "gpu_pipeline_streams",
              "gpu_pipeline_inflight_limit", "gpu_pipeline_prefetch", "gpu_pipeline_align_hop", "gpu_pipeline_config",
              "gpu_pipeline_failures"):
        assert k in meta, k
    assert meta["gpu_pipeline_chunks"] == 4 and meta["gpu_pipeline_used"] is False  # no s_sep stream -> not enabled
    with ctx.acquire_inflight():
        pass


def test_c_abi_library_exports_every_declared_symbol():
    from audio_cut_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "audiocut_b200.h")).read()
    declared = set(re.findall(r"AC_API [^;(]*?\b(ac_\w+)\(", hdr))
    assert len(declared) >= 20
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.EXPORTS)
    assert lib.ac_abi_version() == 1
    assert lib.ac_frame_count(10000, 2048, 441, 1) == 1 + 10000 // 441


def test_ops_fail_loudly_without_cuda():
    import torch

    from audio_cut_b200 import _lib, ops

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(_lib.AudioCutError):
        ops.frame_rms(torch.zeros(1000), 2048, 441)


def test_peak_pick_vectorised_matches_loop_oracle():
    from audio_cut_b200 import host_dsp

    rng = np.random.default_rng(0)
    for n in (1, 5, 201, 1000):
        env = np.abs(rng.standard_normal(n)).astype(np.float32)
        env[rng.random(n) < 0.2] = 0
        for sr, hop in ((44100, 2205), (44100, 512), (44100, 441)):
            np.testing.assert_array_equal(host_dsp.onset_detect(env, sr, hop), OF.onset_detect(env, sr, hop))
    assert host_dsp.onset_detect(np.zeros(10, np.float32), 44100, 512).size == 0


def test_tempo_and_beats_on_click_envelope():
    from audio_cut_b200 import host_dsp

    sr, hop = 44100, 512
    n = int(60 * sr / hop)
    env = np.zeros(n, np.float32)
    period = 60.0 / 100.0 * sr / hop  # 100 BPM
    for k in range(int(n / period)):
        env[int(round(k * period))] = 1.0
    env += 0.01 * np.abs(np.random.default_rng(1).standard_normal(n)).astype(np.float32)
    bpm, beats = host_dsp.beat_track(env, sr, hop)
    assert abs(bpm - 100.0) < 3.0
    iv = np.diff(beats)
    assert len(beats) > 80 and abs(np.median(iv) - period) <= 1.0
    curve = host_dsp.tempo(env, sr, hop, aggregate=None)
    assert curve.shape == (n,) and abs(np.median(curve) - 100.0) < 3.0


def test_unet_blob_layout_sizes():
    import ctypes as C

    from audio_cut_b200 import _lib, unet_weights as uw

    lib = _lib.load()
    for geo in (uw.UNetGeometry(), uw.UNetGeometry(256, 32, 4, 16), uw.UNetGeometry(512, 64, 4, 32)):
        g = _lib.UNetGeom(geo.dim_f, geo.dim_t, geo.dim_c, geo.g, geo.n, geo.l, geo.bn)
        want = lib.ac_unet_param_floats(C.byref(g))
        shapes = uw.param_shapes(geo)
        # blob = all weights + 2 floats (scale, shift) per BatchNorm channel + final bias; conv biases fold away
        n_w = sum(int(np.prod(s)) for k, s in shapes.items() if len(s) > 1)
        n_bn = sum(s[0] for k, s in shapes.items() if k.endswith("running_mean"))
        assert want == n_w + 2 * n_bn + geo.dim_c
    assert lib.ac_unet_param_floats(C.byref(_lib.UNetGeom(100, 32, 4, 16, 5, 3, 8))) == 0  # invalid geometry


def test_oracle_net_matches_folded_numpy_first_layer():
    """fold_bn used by the packer agrees with torch BatchNorm in eval mode."""
    import torch

    from audio_cut_b200 import unet_weights as uw
    from oracle import unet as ounet

    geo = uw.UNetGeometry(256, 32, 4, 16)
    st = uw.random_state(geo)
    net = ounet.build_net(st, 256, 32, 16)
    x = torch.randn(1, 4, 256, 32)
    with torch.no_grad():
        ref = net.first_conv(x).numpy()
    scale, shift = uw.fold_bn(st, "first_conv.1", st["first_conv.0.bias"])
    y = np.einsum("oc,bcft->boft", st["first_conv.0.weight"][:, :, 0, 0], x.numpy())
    y = np.maximum(y * scale[None, :, None, None] + shift[None, :, None, None], 0)
    np.testing.assert_allclose(y, ref, rtol=1e-4, atol=1e-5)


def test_cpp_beat_dp_matches_numpy_loop():
    from audio_cut_b200 import host_dsp, ops

    rng = np.random.default_rng(5)
    for n, period in ((500, 43), (2000, 52), (300, 1), (64, 7)):
        ls = np.abs(rng.standard_normal(n)).astype(np.float32)
        ls[: n // 10] = 0
        b0, c0 = host_dsp._beat_dp(ls, period, 100.0)
        b1, c1 = ops.host_beat_dp(ls, period, 100.0)
        np.testing.assert_array_equal(b0, b1)
        np.testing.assert_allclose(c0, c1, rtol=1e-6, atol=1e-6)


def test_pyin_oracle_recovers_a_harmonic_tone():
    """The librosa.pyin restatement (oracle/features.py): a 220 Hz harmonic tone decodes to the 220 Hz pitch bin."""
    from oracle import features as OF

    sr = 44100
    t = np.arange(int(1.0 * sr)) / sr
    y = (0.5 * np.sin(2 * np.pi * 220 * t) + 0.25 * np.sin(2 * np.pi * 440 * t)).astype(np.float32)
    y[: int(0.25 * sr)] = 0.0
    y += (1e-3 * np.random.default_rng(0).standard_normal(len(y))).astype(np.float32)
    f0, flag, vp = OF.pyin(y, sr, hop_length=441)
    assert f0.shape == (1 + len(y) // 441,)
    assert not flag[:15].any() and flag[40:90].all()
    assert np.allclose(f0[40:90], 220.0, rtol=0.006)  # one 0.1-semitone bin
    W, pmin, pmax, bps, nb = OF.pyin_geometry(sr)
    assert (W, pmin, pmax, bps, nb) == (1024, 21, 675, 10, 601)
    T = OF.transition_local_triangle(nb, 41)
    assert np.allclose(T.sum(axis=1), 1.0) and T[300, 279] == 0 and T[300, 280] > 0 and T[0, 21] == 0


def test_lpc_burg_recovers_an_ar2_process_and_ragged_tracks():
    from audio_cut_b200 import ops
    from oracle import features as OF

    rng = np.random.default_rng(1)
    e = rng.standard_normal(20000)
    y = np.zeros_like(e)
    for i in range(2, len(e)):
        y[i] = 1.3 * y[i - 1] - 0.6 * y[i - 2] + e[i]
    a = OF.lpc_burg(y, 2)
    assert np.allclose(a, [1.0, -1.3, 0.6], atol=0.03)
    mags = np.array([[3.0, 2.0, 1.0], [5.0, 0.0, 0.0], [0.0, 0.0, 0.0], [7.0, 6.0, 0.0]], dtype=np.float32)
    counts = np.array([3, 1, 0, 2])
    t1, t2, t3 = ops.formant_tracks(mags, counts)
    assert t1.tolist() == [3.0, 5.0, 0.0, 7.0] and t2.tolist() == [2.0, 0.0, 6.0] and t3.tolist() == [1.0, 0.0]


def test_public_header_is_plain_c():
    """The drop-in boundary is a C ABI: include/audiocut_b200.h must compile as C99 (and as C++) on its own."""
    import shutil
    import subprocess

    hdr = os.path.join(ROOT, "include", "audiocut_b200.h")
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    subprocess.run(["gcc", "-fsyntax-only", "-std=c99", "-Wall", "-Werror", "-x", "c", hdr], check=True)
    subprocess.run(["g++", "-fsyntax-only", "-std=c++17", "-x", "c++", hdr], check=True)


def test_refine_host_logic_matches_oracle():
    """Host half of the finalize_cut_points drop-in (score-ordered NMS with per-window cap, boundary / gap filter)
    against the oracle restatement that is pinned to the reference's refine.py (no GPU needed for this part)."""
    import importlib.util, os, sys, types

    # audio_cut_b200.refine imports .ops (ctypes binding only; nothing is called here)
    from audio_cut_b200 import refine as R
    from oracle import cuts

    rng = np.random.default_rng(3)
    for case in range(20):
        n = int(rng.integers(1, 40))
        pts = [(float(t), float(s)) for t, s in zip(rng.uniform(0, 60, n), rng.choice([0.1, 0.5, 0.5, 0.9, 0.3], n))]
        gap = float(rng.choice([0.2, 1.0, 2.5]))
        cap = [None, 1, 3][case % 3]
        topk = [None, 5][case % 2]
        win = float(rng.choice([5.0, 10.0]))
        ref = cuts.nms_min_gap(pts, gap, topk=topk, max_per_window=cap, window_s=win)
        got = R.nms_min_gap([R.CutPoint(t, s) for t, s in pts], gap, topk, max_per_window=cap, window_s=win)
        assert [(p.t, p.score) for p in got] == [tuple(p) for p in ref]
    times = [0.2, 0.49, 0.5, 0.51, 3.0, 3.4, 5.0, 9.6, 9.5, 9.51]
    assert R._filter_cut_times(times, duration_s=10.0, min_gap_s=1.0, min_boundary_s=0.5) == [0.51, 3.0, 5.0]
    assert R._filter_cut_times(times, duration_s=0.0, min_gap_s=1.0, min_boundary_s=0.5) == []


# --------------------------------------------------------------------------- model files / drop-in hygiene
def _export_onnx(net, x, path):
    """torch's TorchScript ONNX exporter writes the protobuf itself; only its onnxscript post-pass imports `onnx`."""
    import importlib
    import warnings

    import torch

    mod = importlib.import_module("torch.onnx._internal.torchscript_exporter.onnx_proto_utils")
    mod._add_onnxscript_fn = lambda proto, opsets: proto
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.onnx.export(net, (x,), path, input_names=["input"], output_names=["output"],
                          dynamic_axes={"input": {0: "batch"}, "output": {0: "batch"}}, dynamo=False, opset_version=13)


def test_onnx_model_file_is_read_like_the_reference_session(tmp_path):
    """The reference hands a .onnx to onnxruntime (backends.py:137-181, 216-253); here the initializers are read straight
    out of the protobuf and mapped onto the TFC-TDF parameter names.  A real ONNX file is produced with torch's exporter
    (BatchNorm folded into the convolutions, as in Kim_Vocal_1.onnx) and must give the same network."""
    import torch

    from audio_cut_b200 import onnx_weights as ow
    from audio_cut_b200 import unet_weights as uw
    from oracle import unet as ounet

    try:
        import importlib

        importlib.import_module("torch.onnx._internal.torchscript_exporter.onnx_proto_utils")
    except Exception:
        pytest.skip("this torch build has no TorchScript ONNX exporter")
    geo = uw.UNetGeometry(dim_f=128, dim_t=32, g=16, n=2, l=2, bn=4)
    st = uw.random_state(geo, gains=[])
    net = ounet.build_net(st, geo.dim_f, geo.dim_t, geo.g, L=2 * geo.n + 1, l=geo.l, bn=geo.bn)
    x = torch.randn(2, 4, geo.dim_f, geo.dim_t, generator=torch.Generator().manual_seed(0))
    path = str(tmp_path / "Kim_Vocal_1.onnx")
    _export_onnx(net, x, path)
    nodes, inits = ow.read_onnx(path)
    assert sum(n["op"] == "Conv" for n in nodes) == 1 + (2 * geo.n + 1) * geo.l + geo.n + 1
    st2, geo2 = ow.load_onnx(path, dim_t=geo.dim_t)
    assert geo2 == geo
    assert set(st2) == set(uw.param_shapes(geo))
    net2 = ounet.build_net(st2, geo.dim_f, geo.dim_t, geo.g, L=2 * geo.n + 1, l=geo.l, bn=geo.bn)
    with torch.no_grad():
        a, b = net(x), net2(x)
    assert float((a - b).abs().max()) < 1e-5 * max(1.0, float(a.abs().max()))
    # the packed blob the C ABI takes is the same function too (BN folded either by the exporter or by pack_blob)
    assert uw.pack_blob(st2, geo).shape == uw.pack_blob(st, geo).shape
    # a file that is not this architecture is rejected loudly
    bad = tmp_path / "bad.onnx"
    bad.write_bytes(b"\\x08\\x07")
    with pytest.raises(ValueError):
        ow.load_onnx(str(bad))


def test_chunk_vad_adapter_matches_reference_semantics():
    """B200ChunkVAD vs a hand-computed timeline: halo clipping, the straddling-left-edge rule, 120 ms merge, focus windows
    (silero_chunk_vad.py:56-116, 118-188)."""
    from audio_cut_b200.chunk_vad import B200ChunkVAD
    from audio_cut_b200.gpu_pipeline import chunk_schedule

    sr = 1000
    plans = chunk_schedule(25.0, chunk_s=10.0, overlap_s=2.5, halo_s=0.5)
    stamps = {0: [{"start": 1000, "end": 2000}, {"start": 9200, "end": 9900}],          # second one crosses eff_end 9.5
              1: [{"start": 100, "end": 900}, {"start": 1900, "end": 1950}, {"start": 5000, "end": 5000}],
              2: [{"start": 0, "end": 300}, {"start": 9900, "end": 10000}]}
    vad = B200ChunkVAD(sr, inference_fn=None)
    calls = []
    for p in plans:
        vad.inference_fn = lambda x, i=p.index: (calls.append(i), stamps.get(i, []))[1]
        vad.process_chunk(p, np.zeros(int((p.end_s - p.start_s) * sr), np.float32), sr)
    segs = vad.finalize()
    # chunk 0: [1,2], [9.2,9.5]; chunk 1 (start 7.5, eff 8.0-17.0): [7.6,8.4] straddles the left edge -> keeps 7.6; [9.4,9.45] merges
    # with [9.2,9.5]; chunk 2 (start 15, eff 15.5-25): [15,15.3] lies before its effective region -> dropped; [24.9,25.0]
    got = [(round(s["start"], 6), round(s["end"], 6)) for s in segs]
    assert got == [(1.0, 2.0), (7.6, 8.4), (9.2, 9.5), (24.9, 25.0)], got
    assert all(abs(s["duration"] - (s["end"] - s["start"])) < 1e-12 for s in segs)
    wins = vad.to_focus_windows(pad_s=0.5)
    assert [(round(a, 6), round(b, 6)) for a, b in wins] == [(0.5, 2.5), (7.1, 10.0), (24.4, 25.0)]
    with pytest.raises(ValueError):
        vad.process_chunk(plans[0], np.zeros(10, np.float32), sr + 1)


def test_separator_without_model_raises_like_the_reference(tmp_path, monkeypatch):
    """enhanced_vocal_separator.py:136-137 / backends.py:154-160: no model file -> the constructor raises.  There is no
    silent random-weight network and the default n_fft is the reference's 6144 (backends.py:264)."""
    from audio_cut_b200.backends import B200Mdx23Backend
    from audio_cut_b200.separator import B200VocalSeparator

    monkeypatch.setenv("MDX23_MODELS_PATH", str(tmp_path))
    monkeypatch.delenv("MDX23_N_FFT", raising=False)
    with pytest.raises(RuntimeError, match="no usable separation backend"):
        B200VocalSeparator(44100)
    be = B200Mdx23Backend(str(tmp_path))
    assert be.geom.n_fft == 6144 and be.align_hop == 4096
    with pytest.raises(FileNotFoundError):
        be.load_model()
    monkeypatch.setenv("MDX23_MODEL_FILENAME", "Kim_Vocal_1.onnx")
    with pytest.raises(FileNotFoundError, match="Kim_Vocal_1.onnx"):
        B200Mdx23Backend(str(tmp_path)).load_model()
