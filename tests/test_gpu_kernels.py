"""Parity of every CUDA kernel against the CPU oracle, through the C ABI (run with -m gpu on a B200)."""
import numpy as np
import pytest
import torch

from helpers import rel_err, sdr_db

pytestmark = pytest.mark.gpu

SR = 44100


@pytest.fixture(scope="module")
def ops():
    from audio_cut_b200 import ops as _ops

    return _ops


@pytest.fixture(scope="module")
def audio():
    from audio_cut_b200 import synth

    return synth.synth_track(21.0, seed=3)


# --------------------------------------------------------------------------- framewise RMS / ZCR
@pytest.mark.parametrize("frame,hop", [(4410, 2205), (1102, 441), (2048, 441), (2205, 882), (2048, 512), (100, 50)])
def test_frame_rms_matches_oracle(ops, audio, frame, hop):
    from oracle import features as OF

    y = audio[0]
    got = ops.frame_rms(torch.from_numpy(y).cuda(), frame, hop).cpu().numpy()
    ref = OF.rms(y, frame, hop)
    assert got.shape == ref.shape
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-7)  # north_star: features within 1e-4 relative


def test_frame_rms_edge_cases(ops):
    from oracle import features as OF

    rng = np.random.default_rng(0)
    for n in (1, 7, 441, 2047, 2048, 2049, 5000):
        y = rng.standard_normal(n).astype(np.float32)
        # unaligned device pointer (slice of a larger tensor)
        buf = torch.zeros(n + 3, device="cuda")
        buf[3:] = torch.from_numpy(y).cuda()
        got = ops.frame_rms(buf[3:], 2048, 441).cpu().numpy()
        np.testing.assert_allclose(got, OF.rms(y, 2048, 441), rtol=1e-4, atol=1e-7)


def test_zero_crossing_rate(ops, audio):
    from oracle import features as OF

    y = audio[0][: 5 * SR].copy()
    y[1000:3000] = 0.0
    got = ops.zero_crossing_rate(torch.from_numpy(y).cuda(), 2048, 441).cpu().numpy()
    np.testing.assert_allclose(got, OF.zero_crossing_rate(y, 2048, 441), rtol=0, atol=1e-7)


# --------------------------------------------------------------------------- MDX STFT / iSTFT
GEOMS = [(512, 128, 224, 32), (1280, 256, 512, 64), (640, 128, 256, 32), (6144, 1024, 3072, 256), (7680, 1024, 3072, 256)]


@pytest.mark.parametrize("n_fft,hop,dim_f,dim_t", GEOMS)
def test_stft_mdx_matches_torch(ops, n_fft, hop, dim_f, dim_t):
    from oracle import mdx

    g = mdx.MdxGeometry(n_fft, hop, dim_f, dim_t)
    B = 2
    x = torch.randn(B, 2, g.chunk_size, generator=torch.Generator().manual_seed(1)) * 0.3
    ref = mdx.stft(x, g)  # [B,4,F,T]
    got = ops.stft_mdx(x.cuda(), ops.mdx_geom(n_fft, hop, dim_f, dim_t))
    got = ops.tfc_to_onnx(got).cpu()
    assert got.shape == ref.shape
    assert rel_err(ref.numpy(), got.numpy()) < 2e-6
    assert sdr_db(ref.numpy(), got.numpy()) > 110


@pytest.mark.parametrize("n_fft,hop,dim_f,dim_t", GEOMS)
def test_istft_mdx_matches_torch(ops, n_fft, hop, dim_f, dim_t):
    from oracle import mdx

    g = mdx.MdxGeometry(n_fft, hop, dim_f, dim_t)
    B = 2
    spec = torch.randn(B, 4, dim_f, dim_t, generator=torch.Generator().manual_seed(2))
    ref = mdx.istft(spec, g)
    got = ops.istft_mdx(ops.onnx_to_tfc(spec.cuda()), ops.mdx_geom(n_fft, hop, dim_f, dim_t)).cpu()
    assert got.shape == ref.shape
    assert rel_err(ref.numpy(), got.numpy()) < 5e-6
    assert sdr_db(ref.numpy(), got.numpy()) > 100


# 16-bit formats of the tensor-core path: (name, AC_* dtype, torch dtype).  fp16 (IEEE half, 11-bit significand) is
# the production format - it is the one that meets north_star's >= 40 dB stem-SDR gate; bf16 (8-bit significand)
# runs through the same kernels and is held to what its rounding noise allows.
H16 = [("fp16", 2, torch.float16), ("bf16", 1, torch.bfloat16)]


@pytest.mark.parametrize("name,dtype,tdtype", H16)
def test_stft_16bit_output(ops, name, dtype, tdtype):
    from oracle import mdx

    g = mdx.MdxGeometry(1280, 256, 512, 64)
    x = torch.randn(1, 2, g.chunk_size, generator=torch.Generator().manual_seed(1)) * 0.3
    ref = mdx.stft(x, g)
    got = ops.stft_mdx(x.cuda(), ops.mdx_geom(1280, 256, 512, 64), dtype=dtype)
    assert got.dtype == tdtype
    got = ops.tfc_to_onnx(got).float().cpu()
    assert sdr_db(ref.numpy(), got.numpy()) > (65 if name == "fp16" else 45)  # rounding of the output only
    back = ops.istft_mdx(ops.stft_mdx(x.cuda(), ops.mdx_geom(1280, 256, 512, 64), dtype=dtype), ops.mdx_geom(1280, 256, 512, 64))
    ref_back = mdx.istft(ref, g)
    assert sdr_db(ref_back.numpy(), back.cpu().numpy()) > (60 if name == "fp16" else 40)


# --------------------------------------------------------------------------- U-Net
def _unet_case(ops, dim_f, dim_t, g, B, seed=0):
    from audio_cut_b200 import unet_weights as uw
    from oracle import unet as ounet

    geo = uw.UNetGeometry(dim_f=dim_f, dim_t=dim_t, g=g)
    st = uw.random_state(geo, seed=1234)
    net = ops.UNet(st, geo)
    ref_net = ounet.build_net(st, dim_f, dim_t, g)
    x = torch.randn(B, 4, dim_f, dim_t, generator=torch.Generator().manual_seed(seed)) * 3.0
    with torch.no_grad():
        ref = ref_net(x)
    return net, x, ref


@pytest.mark.parametrize("dim_f,dim_t,g,B", [(256, 32, 16, 2), (512, 64, 32, 1), (256, 32, 48, 3)])
def test_unet_fp32_matches_oracle(ops, dim_f, dim_t, g, B):
    net, x, ref = _unet_case(ops, dim_f, dim_t, g, B)
    got = ops.tfc_to_onnx(net.forward(ops.onnx_to_tfc(x.cuda()))).cpu()
    assert got.shape == ref.shape
    s = sdr_db(ref.numpy(), got.numpy())
    assert s > 80, s  # north_star asks >= 60 dB on stems for the fp32 path


@pytest.mark.parametrize("name,dtype,tdtype", H16)
@pytest.mark.parametrize("dim_f,dim_t,g,B", [(256, 32, 16, 2), (512, 64, 48, 1)])
def test_unet_16bit_simt_close_to_oracle(ops, dim_f, dim_t, g, B, name, dtype, tdtype):
    net, x, ref = _unet_case(ops, dim_f, dim_t, g, B)
    net.set_debug(True)
    got = ops.tfc_to_onnx(net.forward(ops.onnx_to_tfc(x.cuda()).to(tdtype))).float().cpu()
    s = sdr_db(ref.numpy(), got.numpy())
    assert s > (50 if name == "fp16" else 32), s


def test_unet_full_geometry_fp32(ops):
    net, x, ref = _unet_case(ops, 3072, 256, 48, 1)
    got = ops.tfc_to_onnx(net.forward(ops.onnx_to_tfc(x.cuda()))).cpu()
    s = sdr_db(ref.numpy(), got.numpy())
    assert s > 80, s


# --------------------------------------------------------------------------- whole-track separation
def _track_case(ops, n_fft, hop, dim_f, dim_t, g, seconds, stereo, sr, chunk_s, overlap_s, halo_s, align_hop, dtype=0,
                output_is_vocal=True):
    from audio_cut_b200 import synth, unet_weights as uw
    from oracle import mdx, pipeline, planner
    from oracle import unet as ounet

    geo = uw.UNetGeometry(dim_f=dim_f, dim_t=dim_t, g=g)
    st = uw.random_state(geo, seed=1234)
    net = ops.UNet(st, geo)
    ref_net = ounet.build_net(st, dim_f, dim_t, g)
    mg = mdx.MdxGeometry(n_fft, hop, dim_f, dim_t)
    audio = synth.synth_track(seconds, sr=sr, seed=5, stereo=stereo)
    total = audio.shape[-1]
    plans = planner.chunk_schedule(total / float(sr), chunk_s, overlap_s, halo_s)
    bounds = [planner.sample_bounds(p, sr, total) for p in plans]
    otype = "vocal" if output_is_vocal else "instrumental"
    ref_v, ref_i = pipeline.separate_track(
        audio, lambda ch: mdx.infer_chunk(ch, ref_net, mg, align_hop=align_hop, output_type=otype), sr=sr, plans=plans)
    mix = torch.from_numpy(audio if stereo else audio[None, :]).cuda()
    v, i, w = ops.separate_track(net, mix, bounds, ops.mdx_geom(n_fft, hop, dim_f, dim_t), align_hop=align_hop,
                                 output_is_vocal=output_is_vocal, dtype=dtype)
    return ref_v, ref_i, v.cpu().numpy(), i.cpu().numpy(), w.cpu().numpy(), bounds


@pytest.mark.parametrize("stereo,output_is_vocal", [(True, True), (False, True), (True, False)])
def test_separate_track_small_geometry(ops, stereo, output_is_vocal):
    # 8 kHz, 2 s chunks: 11 chunks with ragged tail, seams of weight 2
    ref_v, ref_i, v, i, w, bounds = _track_case(ops, 640, 128, 256, 32, 16, 17.3, stereo, 8000, 2.0, 0.5, 0.1, 256,
                                               output_is_vocal=output_is_vocal)
    assert sdr_db(ref_v, v) > 60 and sdr_db(ref_i, i) > 60, (sdr_db(ref_v, v), sdr_db(ref_i, i))
    wref = np.zeros_like(w)
    for cs, ce, es, ee in bounds:
        wref[es:ee] += 1
    np.testing.assert_array_equal(w, wref)
    assert w.max() == 2 and w.min() == 1


@pytest.mark.parametrize("n_samples,max_batch", [(138400, 4), (138400, 0), (641, 0), (16001, 3)])
def test_separate_track_pipelined_copies_equal_the_plain_call(ops, n_samples, max_batch):
    """ac_separate_track_pipelined (mix uploaded in pieces ahead of the batch that reads them, every finished stretch of
    the stems finalised and downloaded while the next batch runs) against ac_separate_track_ex on a resident mix with one
    download at the end: same kernels on the same data, so the stems must be bit-identical - with several batches per track
    (max_batch 3 / 4: 11+ windows), one batch, and tracks shorter than a window / a chunk."""
    import torch

    from audio_cut_b200 import unet_weights as uw
    from oracle import planner

    sr, n_fft, hop, dim_f, dim_t, g = 8000, 640, 128, 256, 32, 16
    geo = uw.UNetGeometry(dim_f=dim_f, dim_t=dim_t, g=g)
    net = ops.UNet(uw.random_state(geo, seed=1234), geo)
    rng = np.random.default_rng(n_samples + max_batch)
    audio = (0.3 * rng.standard_normal((2, n_samples))).astype(np.float32)
    plans = planner.chunk_schedule(n_samples / float(sr), 2.0, 0.5, 0.1)
    bounds = [planner.sample_bounds(p, sr, n_samples) for p in plans]
    geom = ops.mdx_geom(n_fft, hop, dim_f, dim_t)
    v0, i0, w0 = ops.separate_track(net, torch.from_numpy(audio).cuda(), bounds, geom, align_hop=256, max_batch=max_batch)
    v0, i0, w0 = v0.cpu(), i0.cpu(), w0.cpu()
    host_mix = torch.from_numpy(audio).pin_memory()
    host_out = torch.full((2, n_samples), float("nan")).pin_memory()
    d_mix = torch.full((2, n_samples), float("nan"), device="cuda")  # filled by the call
    copy_stream, up = torch.cuda.Stream(), torch.cuda.Event()
    copy_stream.wait_stream(torch.cuda.current_stream())
    v1, i1, w1 = ops.separate_track(net, d_mix, bounds, geom, align_hop=256, max_batch=max_batch, host_mix=host_mix.data_ptr(),
                                    host_out=(host_out[0].data_ptr(), host_out[1].data_ptr()), copy_stream=copy_stream,
                                    uploaded_event=up)
    up.synchronize()
    assert torch.equal(d_mix.cpu(), torch.from_numpy(audio))  # the whole mix is resident once the event has fired
    torch.cuda.synchronize()
    assert torch.equal(v1.cpu(), v0) and torch.equal(i1.cpu(), i0) and torch.equal(w1.cpu(), w0)
    assert torch.equal(host_out[0], v0) and torch.equal(host_out[1], i0)


@pytest.mark.parametrize("name,dtype,tdtype", H16)
def test_fused_forms_are_bit_identical_to_the_unfused_kernels(ops, name, dtype, tdtype):
    """Kim_Vocal geometry, 12 s stereo (3 windows): the production schedule (level-0 conv chains as one CTA-pair launch each, the
    final 1x1 conv inside the last TDF2, the first 1x1 conv inside the STFT epilogue) against debug mode 3 = the same kernels
    without those fusions.  Each fusion keeps the arithmetic and the 16-bit rounding points of what it replaces, so the stems
    and the weights must be identical bit for bit."""
    import torch

    from audio_cut_b200 import synth, unet_weights as uw
    from oracle import planner

    sr = 44100
    geo = uw.UNetGeometry()
    net = ops.UNet(uw.random_state(geo, seed=1234), geo)
    audio = synth.synth_track(12.0, sr=sr, seed=5, stereo=True)
    n = audio.shape[-1]
    plans = planner.chunk_schedule(n / float(sr), 10.0, 2.5, 0.5)
    bounds = [planner.sample_bounds(p, sr, n) for p in plans]
    geom = ops.mdx_geom(7680, 1024, 3072, 256)
    mix = torch.from_numpy(audio).cuda()
    out = []
    for mode in (0, 3):
        net.set_debug(mode)
        v, i, w = ops.separate_track(net, mix, bounds, geom, dtype=dtype)
        assert _tc_aborted() == 0
        out.append((v.cpu(), i.cpu(), w.cpu()))
    net.set_debug(0)
    assert float(out[0][0].abs().max()) > 1e-3
    for a, b in zip(out[0], out[1]):
        assert torch.equal(a, b), float((a - b).abs().max())


@pytest.mark.parametrize("n_samples", [1, 100, 641, 5000, 16001])
def test_separate_track_short_and_ragged_inputs(ops, n_samples):
    """Tracks shorter than one model window / one chunk / not a multiple of anything: the planner returns a single
    plan (gpu_pipeline.py:341-343), the backend pads to one window (backends.py:306-330)."""
    import torch

    from audio_cut_b200 import unet_weights as uw
    from oracle import mdx, pipeline, planner
    from oracle import unet as ounet

    sr, n_fft, hop, dim_f, dim_t, g = 8000, 640, 128, 256, 32, 16
    geo = uw.UNetGeometry(dim_f=dim_f, dim_t=dim_t, g=g)
    st = uw.random_state(geo, seed=1234)
    net = ops.UNet(st, geo)
    ref_net = ounet.build_net(st, dim_f, dim_t, g)
    mg = mdx.MdxGeometry(n_fft, hop, dim_f, dim_t)
    rng = np.random.default_rng(n_samples)
    audio = (0.3 * rng.standard_normal((2, n_samples))).astype(np.float32)
    plans = planner.chunk_schedule(n_samples / float(sr), 2.0, 0.5, 0.1)
    bounds = [planner.sample_bounds(p, sr, n_samples) for p in plans]
    ref_v, ref_i = pipeline.separate_track(audio, lambda ch: mdx.infer_chunk(ch, ref_net, mg, align_hop=256), sr=sr, plans=plans)
    v, i, w = ops.separate_track(net, torch.from_numpy(audio).cuda(), bounds, ops.mdx_geom(n_fft, hop, dim_f, dim_t), align_hop=256)
    v, i, w = v.cpu().numpy(), i.cpu().numpy(), w.cpu().numpy()
    assert v.shape == ref_v.shape == (n_samples,)
    assert w.min() >= 1
    np.testing.assert_allclose(v, ref_v, atol=2e-4 * max(1e-6, np.abs(ref_v).max()) + 1e-7)
    np.testing.assert_allclose(i, ref_i, atol=2e-4 * max(1e-6, np.abs(ref_i).max()) + 1e-7)


def test_separate_track_full_geometry_30s_stereo(ops):
    """BASELINE config 1: 30 s stereo, Kim_Vocal geometry (n_fft 7680), 4 chunks / 8 windows."""
    ref_v, ref_i, v, i, w, bounds = _track_case(ops, 7680, 1024, 3072, 256, 48, 30.0, True, 44100, 10.0, 2.5, 0.5, 4096)
    assert len(bounds) == 4
    sv, si = sdr_db(ref_v, v), sdr_db(ref_i, i)
    assert sv > 60 and si > 60, (sv, si)


# north_star / BASELINE.md: separated-stem SDR >= 40 dB for the 16-bit tensor-core path (>= 60 dB for fp32), both stems,
# against the CPU oracle on the same input at FULL geometry.  fp16 is the format that carries the gate (measured 48.5 dB
# vocal / 69.8 dB instrumental on this track); bf16's 8-bit significand gives 30.8 / 52.1 dB on the same random-init
# network whichever points are kept in fp32 (profiles/r02_bf16_emulation.md) - it is kept as an option and held
# to a regression floor only.
STEM_GATE_DB = 40.0


@pytest.mark.parametrize("n_fft", [7680, 6144])
def test_16bit_stem_sdr_gate_30s_full_geometry(ops, n_fft):
    """BASELINE configs[0] (30 s stereo, 4 chunks / 8 windows) on the tcgen05 path, both n_fft the survey names."""
    ref_v, ref_i, v, i, w, bounds = _track_case(ops, n_fft, 1024, 3072, 256, 48, 30.0, True, 44100, 10.0, 2.5, 0.5, 4096, dtype=2)
    assert _tc_aborted() == 0
    sv, si = sdr_db(ref_v, v), sdr_db(ref_i, i)
    print(f"fp16 stems n_fft={n_fft}: vocal {sv:.1f} dB, instrumental {si:.1f} dB")
    assert sv >= STEM_GATE_DB and si >= STEM_GATE_DB, (sv, si)


def test_bf16_stem_sdr_floor_30s_full_geometry(ops):
    ref_v, ref_i, v, i, w, bounds = _track_case(ops, 7680, 1024, 3072, 256, 48, 30.0, True, 44100, 10.0, 2.5, 0.5, 4096, dtype=1)
    assert _tc_aborted() == 0
    sv, si = sdr_db(ref_v, v), sdr_db(ref_i, i)
    print(f"bf16 stems: vocal {sv:.1f} dB, instrumental {si:.1f} dB")
    assert sv > 28 and si > 48, (sv, si)


def test_16bit_stem_sdr_gate_4min_full_geometry(ops):
    """BASELINE configs[1]: the benchmarked workload itself (4-min stereo track, 32 chunks / 64 windows, n_fft 7680) on
    the benchmarked fp16 tcgen05 path against the oracle's full CPU pass (about a minute on the box's cores)."""
    ref_v, ref_i, v, i, w, bounds = _track_case(ops, 7680, 1024, 3072, 256, 48, 240.0, True, 44100, 10.0, 2.5, 0.5, 4096, dtype=2)
    assert len(bounds) == 32 and _tc_aborted() == 0
    sv, si = sdr_db(ref_v, v), sdr_db(ref_i, i)
    print(f"fp16 stems 4 min: vocal {sv:.1f} dB, instrumental {si:.1f} dB")
    assert sv >= STEM_GATE_DB and si >= STEM_GATE_DB, (sv, si)


@pytest.mark.parametrize("world,dtype_name", [(2, "f32"), (3, "bf16"), (3, "fp16")])
def test_chunk_sharded_track_equals_single_gpu_stitch(ops, world, dtype_name):
    """BASELINE configs[4]: every rank separates a contiguous block of chunks of ONE track (halo recomputed
    locally, only its own samples uploaded); the host merge of the shards is sample-identical to the
    single-GPU result.  The ranks are emulated one after the other on this GPU."""
    from audio_cut_b200 import sharding, synth, unet_weights as uw
    from audio_cut_b200._lib import AC_BF16, AC_F16, AC_F32
    from oracle import planner

    sr, n_fft, hop, dim_f, dim_t, g = 8000, 640, 128, 256, 32, 16
    geo = uw.UNetGeometry(dim_f=dim_f, dim_t=dim_t, g=g)
    net = ops.UNet(uw.random_state(geo, seed=1234), geo)
    geom = ops.mdx_geom(n_fft, hop, dim_f, dim_t)
    audio = synth.synth_track(23.7, sr=sr, seed=9, stereo=True)
    total = audio.shape[-1]
    plans = planner.chunk_schedule(total / float(sr), 2.0, 0.5, 0.1)
    bounds = [planner.sample_bounds(p, sr, total) for p in plans]
    dtype = {"f32": AC_F32, "bf16": AC_BF16, "fp16": AC_F16}[dtype_name]
    v, i, w = ops.separate_track(net, torch.from_numpy(audio).cuda(), bounds, geom, align_hop=256, dtype=dtype)
    shards = [sharding.separate_chunk_shard(net, geom, audio, bounds, r, world, align_hop=256, dtype=dtype) for r in range(world)]
    assert sum(len(s["weight"]) for s in shards) > total  # neighbouring ranks overlap at the seams
    mv, mi = sharding.merge_chunk_shards(total, shards)
    np.testing.assert_array_equal(mv, v.cpu().numpy())
    np.testing.assert_array_equal(mi, i.cpu().numpy())


def test_full_size_track_properties(ops):
    """BASELINE configs[1] at full size (4-min stereo, Kim_Vocal geometry, fp16 tensor-core path), checked through
    size-independent properties: stem arithmetic is linear (vocal + instrumental == mono mix, backends.py:389-406
    and the uniform overlap average of enhanced_vocal_separator.py:423-458), the overlap weights equal the
    planner's effective-region counts, and the run is deterministic."""
    from audio_cut_b200 import synth, unet_weights as uw
    from audio_cut_b200._lib import AC_F16
    from oracle import planner

    sr = 44100
    geo = uw.UNetGeometry()
    net = ops.UNet(uw.random_state(geo, seed=1234), geo)
    audio = synth.synth_track(240.0, sr=sr, seed=0, stereo=True)
    total = audio.shape[-1]
    plans = planner.chunk_schedule(total / float(sr), 10.0, 2.5, 0.5)
    bounds = [planner.sample_bounds(p, sr, total) for p in plans]
    assert len(bounds) == 32
    mix = torch.from_numpy(audio).cuda()
    geom = ops.mdx_geom(7680, 1024, 3072, 256)
    v, i, w = ops.separate_track(net, mix, bounds, geom, dtype=AC_F16)
    v2, i2, _ = ops.separate_track(net, mix, bounds, geom, dtype=AC_F16)
    assert _tc_aborted() == 0
    assert torch.equal(v, v2) and torch.equal(i, i2)
    wref = np.zeros(total, np.float32)
    for cs, ce, es, ee in bounds:
        wref[es:ee] += 1
    np.testing.assert_array_equal(w.cpu().numpy(), wref)
    mono = mix.mean(dim=0)
    err = (v + i - mono).abs().max().item()
    assert err < 1e-5 * max(1.0, mono.abs().max().item()), err
    assert torch.isfinite(v).all() and float(v.abs().max()) > 0


# --------------------------------------------------------------------------- STFT-2048 features
@pytest.mark.parametrize("hop", [2205, 441, 512])
def test_stft_features_match_oracle(ops, audio, hop):
    from oracle import features as OF

    y = audio[0]
    # two independent segments (chunks): the top_db clip reference is per segment
    segs = [(0, 10 * SR), (7 * SR + 11, 9 * SR + 5)]
    offs, total = [], 0
    for s, l in segs:
        offs.append(total)
        total += 1 + l // hop
    out = ops.stft_features(torch.from_numpy(y).cuda(), [(s, l, o) for (s, l), o in zip(segs, offs)], hop, SR,
                            total_frames=total, want=("flatness", "onset_mean", "onset_median", "centroid", "low_ratio"))
    out = {k: v.cpu().numpy() for k, v in out.items()}
    for (s, l), o in zip(segs, offs):
        seg = y[s : s + l]
        n = 1 + l // hop
        sl = slice(o, o + n)
        np.testing.assert_allclose(out["flatness"][sl], OF.spectral_flatness(seg, 2048, hop), rtol=1e-4, atol=1e-9)
        ref_mean = OF.onset_strength(seg, SR, hop, aggregate=np.mean)
        ref_med = OF.onset_strength(seg, SR, hop, aggregate=np.median)
        np.testing.assert_allclose(out["onset_mean"][sl], ref_mean, rtol=1e-4, atol=1e-4 * ref_mean.max())
        np.testing.assert_allclose(out["onset_median"][sl], ref_med, rtol=1e-4, atol=1e-4 * max(ref_med.max(), 1e-3))
        np.testing.assert_allclose(out["centroid"][sl], OF.spectral_centroid(seg, SR, 2048, hop), rtol=1e-4)
        np.testing.assert_allclose(out["low_ratio"][sl], OF.low_band_ratio(seg, 2048, hop), rtol=1e-4, atol=1e-7)


# --------------------------------------------------------------------------- tcgen05 kernels
def _tc_aborted():
    from audio_cut_b200 import _lib

    return _lib.load().ac_debug_tc_aborted()


@pytest.mark.parametrize("name,dtype,tdtype", H16)
@pytest.mark.parametrize("dim_f,dim_t,g,B", [(256, 32, 16, 2), (512, 64, 48, 1), (512, 32, 32, 2)])
def test_unet_tcgen05_matches_simt(ops, dim_f, dim_t, g, B, name, dtype, tdtype):
    """Same 16-bit storage, same fp32 accumulation: tensor-core layers vs the CUDA-core kernels."""
    net, x, ref = _unet_case(ops, dim_f, dim_t, g, B)
    xin = ops.onnx_to_tfc(x.cuda()).to(tdtype)
    net.set_debug(True)
    simt = net.forward(xin).float().cpu().numpy()
    net.set_debug(False)
    tc = net.forward(xin).float().cpu().numpy()
    assert _tc_aborted() == 0, "a tcgen05 kernel hit its mbarrier watchdog"
    s = sdr_db(simt, tc)
    assert s > (58 if name == "fp16" else 40), s
    s_ref = sdr_db(ops.onnx_to_tfc(ref).numpy(), tc)
    assert s_ref > (50 if name == "fp16" else 32), s_ref


@pytest.mark.parametrize("name,dtype,tdtype", H16)
@pytest.mark.parametrize("C,T,F,impls", [(48, 24, 640, (2, 3)), (96, 16, 512, (2, 3)), (144, 12, 384, (1, 4)), (192, 8, 256, (1, 4))])
def test_conv3x3_variants_match_cuda_core_layer(ops, C, T, F, impls, name, dtype, tdtype):
    """Layer-level check of every tcgen05 3x3-conv kernel against the CUDA-core implicit GEMM (same 16-bit
    inputs, fp32 accumulation): 1 streaming, 2 weight-stationary, 3 row-stacked (C=48) / CTA pair (C=96),
    4 CTA-pair streaming."""
    rng = np.random.default_rng(C)
    x = torch.randn(3, T, F, C, device="cuda").to(tdtype)
    w = (rng.standard_normal((C, C, 3, 3)) / np.sqrt(9 * C)).astype(np.float32)
    scale = torch.rand(C, device="cuda") + 0.5
    shift = torch.randn(C, device="cuda") * 0.1
    ref, _ = ops.debug_conv3x3(x, w, scale, shift, 0)
    for impl in impls:
        y, _ = ops.debug_conv3x3(x, w, scale, shift, impl)
        assert _tc_aborted() == 0
        s = sdr_db(ref.float().cpu().numpy(), y.float().cpu().numpy())
        assert s > (75 if name == "fp16" else 60), (impl, s)


@pytest.mark.parametrize("name,dtype,tdtype", H16)
@pytest.mark.parametrize("B,T,F", [(2, 24, 640), (1, 7, 130), (3, 40, 3072)])
def test_fused_conv_chain_is_bit_identical_to_three_launches(ops, B, T, F, name, dtype, tdtype):
    """The level-0 TFC conv chain (3 x conv3x3 + BN + ReLU, C = 48) in one kernel with both intermediates in shared
    memory (unet_tc_conv_f3.cu) vs three launches of the weight-stationary kernel: same accumulation order per element,
    same 16-bit rounding of the intermediates -> identical bits.  Shapes: several strips with a ragged last one, a single
    130-position strip with segments shorter than the pipeline depth, and the full frequency axis."""
    C = 48
    rng = np.random.default_rng(B * 1000 + T)
    x = torch.randn(B, T, F, C, device="cuda").to(tdtype)
    w = (rng.standard_normal((3, C, C, 3, 3)) / np.sqrt(9 * C) * 1.6).astype(np.float32)
    scale = torch.rand(3, C, device="cuda") + 0.5
    shift = torch.randn(3, C, device="cuda") * 0.1
    ref, _ = ops.debug_conv3x3_chain(x, w, scale, shift, 0)
    assert _tc_aborted() == 0
    y, _ = ops.debug_conv3x3_chain(x, w, scale, shift, 1)
    assert _tc_aborted() == 0
    assert float(ref.float().abs().max()) > 0.1
    assert torch.equal(ref.view(torch.int16), y.view(torch.int16)), (
        int((ref.view(torch.int16) != y.view(torch.int16)).sum()), sdr_db(ref.float().cpu().numpy(), y.float().cpu().numpy()))
    # and against the CUDA-core implicit GEMM, layer by layer (independent implementation)
    z = x
    for j in range(3):
        z, _ = ops.debug_conv3x3(z, w[j], scale[j], shift[j], 0)
    s = sdr_db(z.float().cpu().numpy(), y.float().cpu().numpy())
    assert s > (70 if name == "fp16" else 50), s


@pytest.mark.parametrize("name,dtype,tdtype", H16)
def test_unet_tcgen05_full_geometry(ops, name, dtype, tdtype):
    """The network alone at Kim_Vocal geometry on a white-noise spectrogram: fp16 63 dB, bf16 45 dB measured."""
    net, x, ref = _unet_case(ops, 3072, 256, 48, 1)
    tc = ops.tfc_to_onnx(net.forward(ops.onnx_to_tfc(x.cuda()).to(tdtype))).float().cpu()
    assert _tc_aborted() == 0
    s = sdr_db(ref.numpy(), tc.numpy())
    assert s > (58 if name == "fp16" else 40), s


def test_tempogram_stats_match_oracle(ops):
    """GPU tempogram / per-frame tempo / global tempo (float32) vs oracle.rhythm (the loop restatement of
    librosa.feature.tempogram / rhythm.tempo, pinned by tests/test_oracle_golden.py) - rows A14 / N2."""
    from oracle import rhythm as R

    for sr, hop, n, bpm_true in ((44100, 512, 5000, 100.0), (44100, 2205, 1203, 128.0)):
        period = 60.0 / bpm_true * sr / hop
        env = np.zeros(n, np.float32)
        for k in range(int(n / period)):
            env[int(round(k * period))] = 1.0
        env += 0.05 * np.abs(np.random.default_rng(2).standard_normal(n)).astype(np.float32)
        win = int(np.floor(8.0 * sr / hop))
        tg = R.tempogram(env, win)
        ref_curve = R.tempo(env, sr, hop, aggregate=None, tg=tg)
        ref_glob = R.tempo(env, sr, hop, tg=tg)[0]
        curve, glob, tg_mean = ops.tempogram_stats(torch.from_numpy(env).cuda(), sr, hop)
        np.testing.assert_allclose(tg_mean, tg.mean(axis=1), rtol=0, atol=2e-5)
        assert glob == ref_glob
        assert np.mean(curve == ref_curve) > 0.995  # float32 vs float64 near-ties between adjacent lags


def test_bpm_features_and_beats_match_oracle(ops):
    """BPMAnalyzer.extract_bpm_features and the cache's tempo curve / beat times through the package (GPU onset envelope and
    tempogram, C++ beat DP) vs oracle.rhythm on the same waveform: same BPM, class, beats, stability, variance, factors."""
    from audio_cut_b200 import synth
    from audio_cut_b200.features_cache import bpm_features_from_wave
    from oracle import features as OF
    from oracle import rhythm as R

    for wave in (synth.synth_track(20.0, seed=2, stereo=False), synth.synth_song(24.0, seed=1)):
        got = bpm_features_from_wave(torch.from_numpy(wave).cuda(), SR)
        ref = R.extract_bpm_features(wave, SR)
        assert got.main_bpm == ref.main_bpm and got.bpm_category == ref.bpm_category
        assert [int(b) for b in got.beat_positions] == [int(b) for b in ref.beat_positions]
        assert abs(got.beat_strength - ref.beat_strength) < 1e-9
        assert abs(got.tempo_variance - ref.tempo_variance) < 2e-3  # per-frame argmax near-ties (float32 tempogram)
        for k, v in ref.adaptive_factors.items():
            g = got.adaptive_factors[k]
            assert (abs(g - v) < 1e-3) if isinstance(v, float) else (g == v), k
        # hop 2205 side (features_cache.py:283-294): beat_track on the cache's own onset envelope
        env = OF.onset_strength(wave, SR, 2205)
        curve, glob, _ = ops.tempogram_stats(torch.from_numpy(env).cuda(), SR, 2205)
        assert glob == R.tempo(env, SR, 2205)[0]
        from audio_cut_b200 import host_dsp

        _, beats = host_dsp.beat_track(env, SR, 2205, bpm=glob, dp=ops.host_beat_dp)
        _, ref_beats = R.beat_track(env, SR, 2205)
        assert list(beats) == list(ref_beats)


# --------------------------------------------------------------------------- pYIN / LPC formants (A17, A18)
def _voiced_test_signal(seconds=3.0, seed=5):
    """Vibrato harmonic tone with pauses + noise floor + an unvoiced noise burst."""
    rng = np.random.default_rng(seed)
    n = int(seconds * SR)
    t = np.arange(n) / SR
    f0 = 180.0 * 2 ** (0.5 * np.sin(2 * np.pi * 0.7 * t)) * (1 + 0.01 * np.sin(2 * np.pi * 5.5 * t))
    ph = 2 * np.pi * np.cumsum(f0) / SR
    y = sum(a * np.sin(k * ph) for k, a in ((1, 0.5), (2, 0.3), (3, 0.15), (4, 0.08)))
    gate = (np.sin(2 * np.pi * 0.9 * t) > -0.3).astype(np.float64)
    y = y * gate + 2e-3 * rng.standard_normal(n)
    a, b = int(0.72 * n), int(0.82 * n)
    y[a:b] = 0.2 * rng.standard_normal(b - a)
    return y.astype(np.float32)


def test_pyin_matches_oracle(ops):
    from oracle import features as OF

    y = _voiced_test_signal()
    f0, flag, vp = ops.pyin(torch.from_numpy(y).cuda(), SR, 441)
    f0, flag, vp = f0.cpu().numpy(), flag.cpu().numpy(), vp.cpu().numpy()
    r_f0, r_flag, r_vp = OF.pyin(y, SR, hop_length=441)
    assert f0.shape == r_f0.shape == (1 + len(y) // 441,)
    # voiced probability is a sum of beta-weighted masses: fp32 CMND vs the fp64 oracle can move a trough
    # across one of the 100 thresholds, which shifts at most one beta weight (<= 0.07) in a frame
    assert np.mean(np.abs(vp - r_vp) < 1e-3) > 0.97, np.mean(np.abs(vp - r_vp) < 1e-3)
    assert np.max(np.abs(vp - r_vp)) < 0.15
    agree = flag == r_flag
    assert agree.mean() > 0.99, agree.mean()
    both = flag & r_flag
    assert both.sum() > 50
    same = np.isclose(f0[both], r_f0[both], rtol=1e-5)
    assert same.mean() > 0.99, same.mean()
    # never more than one 0.1-semitone bin apart
    assert np.max(np.abs(np.log2(f0[both] / r_f0[both]))) * 120 < 1.01


def test_pyin_candidate_stage_only(ops):
    y = _voiced_test_signal(1.0)
    f0, flag, vp = ops.pyin(torch.from_numpy(y).cuda(), SR, 441, decode=False)
    assert f0 is None and flag is None and vp.shape == (1 + len(y) // 441,)
    assert float(vp.max()) <= 1.0 and float(vp.min()) >= 0.0


def test_lpc_formants_match_oracle(ops):
    from oracle import features as OF

    y = _voiced_test_signal(2.0)
    y[5000:9000] = 0.0  # all-zero frames: the reference's "no peaks" branch
    mags, counts = ops.lpc_formants(torch.from_numpy(y).cuda(), SR, 441, 12)
    mags, counts = mags.cpu().numpy(), counts.cpu().numpy()
    r_mags, r_counts = OF.lpc_formant_frames(y, SR, int(0.025 * SR), 441, 12)
    assert mags.shape == r_mags.shape == (len(range(0, len(y) - int(0.025 * SR), 441)), 3)
    same = counts == r_counts
    print("lpc: peak counts equal in", same.mean(), "of", same.size, "frames; max rel magnitude error",
          float(np.max(np.abs(mags[same] - r_mags[same]) / np.maximum(np.abs(r_mags[same]), 1e-12))))
    # Burg recursion AND the 512-point response in fp64 on both sides (north_star: features within 1e-4 relative)
    assert same.all(), np.nonzero(~same)[0][:10]
    np.testing.assert_allclose(mags, r_mags, rtol=1e-4, atol=1e-7)
    # ragged track reconstruction follows the reference's append rule
    tr = ops.formant_tracks(mags, counts)
    assert len(tr[0]) >= len(tr[1]) >= len(tr[2])
    assert len(tr[0]) == int(np.sum((counts > 0) | (counts == 0)))


# --------------------------------------------------------------------------- cut-point refinement
def _refine_on_gpu(mix, vocal, sr, pts, kw, as_tensor=False):
    from audio_cut_b200 import refine as R

    if as_tensor:
        mix = torch.from_numpy(mix).cuda()
        vocal = None if vocal is None else torch.from_numpy(vocal).cuda()
    return R.finalize_cut_points(R.CutContext(sr=sr, mix_wave=mix, vocal_wave=vocal), [R.CutPoint(t=a, score=b) for a, b in pts], **kw)


def test_refine_cut_points_matches_reference_fixture_44k():
    """audio_cut_b200.refine.finalize_cut_points vs the REFERENCE's refine.py (tests/golden/cuts_44k.json): sample
    boundaries, final times and every kept adjustment (raw / guard / final) bit-exact - integer and fp64 outputs."""
    import json, os
    from helpers import cut_case

    cases = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "cuts_44k.json")))
    for case in cases:
        mix, vocal, sr, pts, kw = cut_case(case["seed"])
        res = _refine_on_gpu(mix, vocal, sr, pts, kw, as_tensor=(case["seed"] % 2 == 1 and mix.ndim == 1))
        assert res.sample_boundaries == case["sample_boundaries"], case["seed"]
        assert [repr(float(p.t)) for p in res.final_points] == case["final_times"], case["seed"]
        got = [[repr(a.raw_time), repr(a.guard_time), repr(a.final_time)] for a in res.adjustments]
        assert got == case["adjustments"], case["seed"]
        assert len(res.suppressed_points) == case["n_suppressed"]


def test_refine_cut_points_matches_reference_fixture_8k():
    """The 8 kHz fixture the oracle is pinned with (tests/golden/cuts.json; 450 ms search, 80 ms guard window)."""
    import json, os

    cases = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "cuts.json")))
    for case in cases:
        seed = case["seed"]
        rng = np.random.default_rng(seed)
        sr = 8000
        n = int(24.0 * sr)
        t = np.arange(n) / sr
        s = seed - 100
        gate = (np.sin(2 * np.pi * 0.23 * t + s) > -0.3).astype(np.float64)
        vocal = (0.3 * np.sin(2 * np.pi * 180 * t) * gate + 1e-4 * rng.standard_normal(n)).astype(np.float32)
        mix = (vocal + 0.1 * np.sin(2 * np.pi * 55 * t) * (np.sin(2 * np.pi * 0.11 * t) > 0) + 1e-3 * rng.standard_normal(n)).astype(np.float32)
        res = _refine_on_gpu(mix, vocal, sr, [tuple(p) for p in case["points"]], case["kwargs"])
        assert res.sample_boundaries == case["sample_boundaries"], seed
        assert [repr(float(p.t)) for p in res.final_points] == case["final_times"], seed


def test_refine_cut_points_full_track_matches_oracle(ops):
    """4-minute stems (BASELINE configs[1] length), 120 candidates, default parameters: every per-point time equals the
    oracle's (fp64, compared exactly); the whole-track lookup array agrees to 1e-9 dB where it is evaluated."""
    from audio_cut_b200 import synth
    from oracle import cuts

    mix2 = synth.synth_track(240.0, seed=11)
    mix = mix2.mean(axis=0).astype(np.float32)
    vocal = (0.6 * mix * (np.sin(2 * np.pi * 0.2 * np.arange(mix.size) / SR) > 0)).astype(np.float32)
    rng = np.random.default_rng(5)
    pts = [(float(a), float(b)) for a, b in zip(rng.uniform(0.3, 239.7, 120), rng.uniform(0, 1, 120))]
    kw = dict(min_gap_s=1.0, topk_per_10s=6, floor_db=-50.0)
    res = _refine_on_gpu(mix, vocal, SR, pts, kw, as_tensor=True)
    bounds, times = cuts.finalize_cut_points(mix, vocal, SR, pts, **kw)
    assert res.sample_boundaries == bounds
    assert [p.t for p in res.final_points] == times
    assert len(times) > 20
    win = 441
    got = ops.quiet_lookup_db(torch.from_numpy(vocal[: 30 * SR]).cuda(), win).cpu().numpy()
    ref, _ = cuts.prepare_quiet_lookup(vocal[: 30 * SR], SR, 10.0, -60.0)
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-9)


def test_refine_cut_points_edge_cases():
    from audio_cut_b200 import refine as R

    z = np.zeros(4410, np.float32)
    res = R.finalize_cut_points(R.CutContext(sr=SR, mix_wave=z), [])
    assert res.sample_boundaries == [0, 4410] and res.final_points == []
    # all-zero track: every window ties exactly, nothing may move; a single sample; a point beyond the end
    res = R.finalize_cut_points(R.CutContext(sr=SR, mix_wave=np.zeros(3 * SR, np.float32), vocal_wave=np.zeros(3 * SR, np.float32)),
                                [R.CutPoint(1.0, 0.5), R.CutPoint(2.2, 0.9), R.CutPoint(7.0, 0.1)], min_boundary_s=0.1)
    from oracle import cuts

    zz = np.zeros(3 * SR, np.float32)
    _, times = cuts.finalize_cut_points(zz, zz, SR, [(1.0, 0.5), (2.2, 0.9), (7.0, 0.1)], min_boundary_s=0.1)
    assert [a.final_time for a in res.adjustments] == times == [1.0, 2.2]
    res = R.finalize_cut_points(R.CutContext(sr=SR, mix_wave=np.ones(1, np.float32)), [R.CutPoint(0.0, 1.0)])
    assert res.sample_boundaries == [0, 1]


def test_pyin_viterbi_kernels_agree_bit_for_bit(ops):
    """The production pYIN decode (4-CTA cluster Viterbi + map-composition backtrack) against the single-CTA
    one-barrier kernel, the tiled and the run-time band-width kernels and the serial backtrack: identical f0 / flags."""
    import os
    from audio_cut_b200 import synth

    y = synth.synth_track(20.0, seed=8, stereo=False).astype(np.float32)
    x = torch.from_numpy(y).cuda()
    outs = {}
    try:
        for mode in ("", "generic", "tiled", "fast", "serial-backtrack"):
            os.environ.pop("AC_PYIN_VITERBI", None)
            os.environ.pop("AC_PYIN_BACKTRACK", None)
            if mode == "serial-backtrack":
                os.environ["AC_PYIN_BACKTRACK"] = "serial"
            elif mode:
                os.environ["AC_PYIN_VITERBI"] = mode
            f0, fl, vp = ops.pyin(x)
            outs[mode] = (f0.cpu().numpy(), fl.cpu().numpy(), vp.cpu().numpy())
    finally:
        os.environ.pop("AC_PYIN_VITERBI", None)
        os.environ.pop("AC_PYIN_BACKTRACK", None)
    assert outs[""][1].sum() > 100
    for mode in ("generic", "tiled", "fast", "serial-backtrack"):
        for a, b in zip(outs[""], outs[mode]):
            np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("n", [300, 441, 900, 1400, 441 * 65 + 7, 441 * 129])
def test_pyin_decode_on_very_short_inputs(ops, n):
    """1, 2, 3, 4 frames and lengths around the 64-step backtrack blocks: the cluster Viterbi + composed backtrack agree
    with the run-time-width kernel + serial backtrack (pipeline prologues, block boundaries, empty tails)."""
    import os

    rng = np.random.default_rng(n)
    t = np.arange(n) / SR
    y = (0.4 * np.sin(2 * np.pi * 220.0 * t) + 0.01 * rng.standard_normal(n)).astype(np.float32)
    x = torch.from_numpy(y).cuda()
    try:
        os.environ.pop("AC_PYIN_VITERBI", None)
        os.environ.pop("AC_PYIN_BACKTRACK", None)
        a = [v.cpu().numpy() for v in ops.pyin(x)]
        os.environ["AC_PYIN_VITERBI"] = "generic"
        os.environ["AC_PYIN_BACKTRACK"] = "serial"
        b = [v.cpu().numpy() for v in ops.pyin(x)]
    finally:
        os.environ.pop("AC_PYIN_VITERBI", None)
        os.environ.pop("AC_PYIN_BACKTRACK", None)
    assert a[0].shape == (1 + n // 441,)
    for u, v in zip(a, b):
        np.testing.assert_array_equal(u, v)


def test_frame_rms_segments_one_launch_equals_per_chunk_calls(ops, audio):
    """ac_frame_rms_segments: every pipeline chunk of a track in one launch == librosa.feature.rms per chunk slice
    (features_cache.py:182), including unaligned chunk starts, ragged tails and more segments than one launch carries."""
    from oracle import features as OF

    y = audio[0]
    x = torch.from_numpy(y).cuda()
    rng = np.random.default_rng(1)
    for frame, hop, n_seg in ((4410, 2205, 5), (1102, 441, 7), (2048, 441, 130)):
        segs, off = [], 0
        for _ in range(n_seg):
            ln = int(rng.integers(1, 60000))
            st = int(rng.integers(0, len(y) - ln))
            segs.append((st, ln, off))
            off += ops.frame_count(ln, frame, hop)
        l0 = ops._lib.load().ac_launch_count()
        got = ops.frame_rms_segments(x, segs, frame, hop, total_frames=off).cpu().numpy()
        assert ops._lib.load().ac_launch_count() - l0 == (n_seg + 95) // 96
        for st, ln, o in segs:
            ref = OF.rms(y[st:st + ln], frame, hop)
            np.testing.assert_allclose(got[o:o + len(ref)], ref, rtol=1e-4, atol=1e-7)
