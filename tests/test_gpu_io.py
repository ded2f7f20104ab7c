"""SURVEY.md section 8(f) N3 / N4 on the GPU: PCM decode / pack bit-exact against the oracle (integer work), the polyphase
resampler against scipy.signal.resample_poly, and the batched per-chunk VAD front end through the separator."""
import wave

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _wave(n, ch, seed):
    rng = np.random.default_rng(seed)
    x = 0.7 * np.sin(2 * np.pi * 220.0 * np.arange(n) / 44100.0)[None, :] * np.linspace(0.2, 1.0, ch)[:, None]
    x = x + 0.05 * rng.standard_normal((ch, n))
    x[:, :5] = [[0.0, 1.0, -1.0, 0.5, -0.5]]
    x[:, 5:9] = np.array([2.5, 3.5, 4194303.5, -2.5]) / 8388607.0  # ties of the float32 product: round half to even
    return x.astype(np.float32)


@pytest.mark.parametrize("ch,n", [(1, 1), (1, 4099), (2, 70001)])
def test_pcm_pack_is_bit_exact(ch, n):
    from audio_cut_b200 import audio_io as A
    from oracle import audio_io as IO

    x = _wave(max(n, 9), ch, n)[:, :n] if n >= 9 else _wave(9, ch, n)[:, :n]
    xd = torch.from_numpy(x).cuda()
    frames_first = np.ascontiguousarray(x.T)
    assert A.pack_pcm(xd if ch > 1 else xd[0], "PCM_24").cpu().numpy().tobytes() == IO.pcm24_bytes(frames_first)
    assert A.pack_pcm(xd, "int16").cpu().numpy().tobytes() == IO.int16_bytes(frames_first)
    loud = (x * 3.0).astype(np.float32)  # beyond full scale: the default path wraps like libsndfile, the clip path saturates
    assert A.pack_pcm(torch.from_numpy(loud).cuda(), "PCM_24").cpu().numpy().tobytes() == IO.pcm24_bytes(np.ascontiguousarray(loud.T))
    assert A.pack_pcm(torch.from_numpy(loud).cuda(), "PCM_24_clip").cpu().numpy().tobytes() == IO.pcm24_bytes(np.ascontiguousarray(loud.T), clip=True)


@pytest.mark.parametrize("bits", [16, 24])
@pytest.mark.parametrize("ch", [1, 2])
def test_pcm_decode_mono_and_peak_normalise_are_bit_exact(bits, ch):
    from audio_cut_b200 import audio_io as A
    from oracle import audio_io as IO

    x = _wave(50021, ch, bits + ch)
    data = IO.pcm24_bytes(np.ascontiguousarray(x.T)) if bits == 24 else IO.int16_bytes(np.ascontiguousarray(x.T))
    raw = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    planar = A.decode_pcm(raw, x.shape[1], ch, bits, mono=False).cpu().numpy()
    np.testing.assert_array_equal(planar.reshape(ch, -1), IO.decode_pcm(data, ch, bits, mono=False))
    mono = A.decode_pcm(raw, x.shape[1], ch, bits, mono=True, normalize=True).cpu().numpy()
    np.testing.assert_array_equal(mono, IO.decode_pcm(data, ch, bits, mono=True, normalize=True))
    z = A.peak_normalize_(torch.zeros(100, device="cuda"))
    assert float(z.abs().max()) == 0.0  # all-zero input is left alone (audio_processor.py:55)


def test_wav_file_round_trip(tmp_path):
    """export_audio -> file -> AudioProcessor.load_audio, both ends on the GPU, the container through Python's wave."""
    from audio_cut_b200 import audio_io as A
    from oracle import audio_io as IO

    x = _wave(44100, 2, 5)
    path = str(tmp_path / "seg.wav")
    A.write_wav(path, torch.from_numpy(x).cuda(), 44100)
    with wave.open(path, "rb") as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (2, 3, 44100, 44100)
        data = w.readframes(44100)
    assert data == IO.pcm24_bytes(np.ascontiguousarray(x.T))
    audio, sr = A.load_wav(path, mono=True, normalize=True)
    assert sr == 44100
    np.testing.assert_array_equal(audio.cpu().numpy(), IO.decode_pcm(data, 2, 24, mono=True, normalize=True))


def test_resample_chunks_matches_scipy_resample_poly():
    from audio_cut_b200 import ops
    from oracle import audio_io as IO

    rng = np.random.default_rng(2)
    lens = [441000, 330750, 1, 441, 100003]
    parts = [(0.3 * np.sin(2 * np.pi * (200 + 50 * k) * np.arange(n) / 44100.0) + 0.05 * rng.standard_normal(n)).astype(np.float32)
             for k, n in enumerate(lens)]
    x = torch.from_numpy(np.concatenate(parts)).cuda()
    l0 = ops._lib.load().ac_launch_count()
    batch, out_lens = ops.resample_chunks(x, lens, 44100, 16000)
    assert ops._lib.load().ac_launch_count() - l0 == 1
    assert batch.shape[0] == len(lens) and batch.shape[1] % 4096 == 0  # silero_length_bucket (vocal_pause_detector.py:190-195)
    got = batch.cpu().numpy()
    for k, p in enumerate(parts):
        ref = IO.resample(p)
        assert out_lens[k] == len(ref)
        np.testing.assert_allclose(got[k, : len(ref)], ref, rtol=0, atol=3e-6 * max(1.0, np.abs(ref).max()))
        assert not got[k, len(ref):].any()


def test_batched_vad_front_end_through_the_separator():
    """N4: the per-chunk VAD of enhanced_vocal_separator.py:412-417 as ONE batch - every chunk's own vocal output, resampled to
    16 kHz on the GPU, handed to a batched model hook; the adapter's timeline logic is the SileroChunkVAD one."""
    from audio_cut_b200 import ops, synth, unet_weights as uw
    from audio_cut_b200.backends import B200Mdx23Backend
    from audio_cut_b200.chunk_vad import B200ChunkVAD
    from audio_cut_b200.gpu_pipeline import PipelineConfig, chunk_schedule
    from oracle import audio_io as IO
    from oracle import mdx
    from oracle import unet as ounet

    sr = 44100
    geo = uw.UNetGeometry(dim_f=256, dim_t=32, g=16)
    st = uw.random_state(geo)
    be = B200Mdx23Backend(weights=st, geometry=geo, n_fft=640, hop=128, align_hop=256, precision="fp32", output_type="vocal")
    be.load_model()
    audio = synth.synth_track(1.3, sr=sr, stereo=False)
    plans = chunk_schedule(len(audio) / float(sr), chunk_s=0.5, overlap_s=0.1, halo_s=0.02)
    bounds = [p.sample_bounds(sr, len(audio)) for p in plans]
    lens = [ce - cs for cs, ce, _, _ in bounds]
    side = torch.empty(sum(lens), dtype=torch.float32, device="cuda")
    mix = torch.from_numpy(audio[None, :]).cuda()
    ops.separate_track(be.net, mix, bounds, be.geom, align_hop=256, chunk_vocal=side)
    seen = {}

    def model(batch, out_lens):  # stands in for the Silero network: "speech" wherever the 16 kHz chunk is non-silent
        seen["batch"], seen["lens"] = batch, out_lens
        return [[{"start": 0, "end": n}] for n in out_lens]

    vad = B200ChunkVAD(sr, batch_inference_fn=model)
    vad.process_track(plans, side, lens)
    segs = vad.finalize()
    assert len(segs) == 1 and abs(segs[0]["start"]) < 1e-9 and abs(segs[0]["end"] - plans[-1].effective_end_s) < 2e-4
    # the batch rows are the chunks' own outputs at 16 kHz
    ref_net = ounet.build_net(st, geo.dim_f, geo.dim_t, geo.g)
    mg = mdx.MdxGeometry(640, 128, 256, 32)
    got = seen["batch"].cpu().numpy()
    for k, (cs, ce, _, _) in enumerate(bounds):
        rv, _ = mdx.infer_chunk(audio[cs:ce], ref_net, mg, align_hop=256)
        ref = IO.resample(rv)
        assert seen["lens"][k] == len(ref)
        np.testing.assert_allclose(got[k, : len(ref)], ref, rtol=0, atol=2e-5 * max(1e-3, np.abs(ref).max()) + 1e-7)
