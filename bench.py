#!/usr/bin/env python
"""Headline benchmark: x-realtime separation+features of the audio-cut hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision fp16|bf16|fp32]

A "step" is one pass of the hot path over one synthetic 4-minute 44.1 kHz stereo track per GPU
(BASELINE.json configs[1]; at N > 1 configs[3]: tracks sharded across ranks, no data-path collective,
weak scaling).  Prints ONE JSON line on rank 0 (see the task contract for the keys).

  value  audio seconds / second with the track already resident in HBM: chunked separation
         (STFT -> TFC-TDF U-Net -> fused iSTFT/OLA/stems) + every framewise series the v2.2_mdd path
         consumes (TrackFeatureCache RMS/flatness/onset per chunk, BPM onset envelope, vocal RMS
         1102/441, 2048/441, 2205/882, vocal flatness @441, mix RMS 2048/441).
  e2e    the same metric through the reference-facing call B200VocalSeparator.separate_for_detection()
         with HOST numpy buffers: pinned H2D of the mix, the kernels, D2H of both stems and all
         series, and the host-side rhythm scans (tempogram / beat DP) of TrackFeatureCache.
  --impl reference   times the CPU restatement of the reference path (oracle/: torch-CPU network +
         numpy features; onnxruntime / librosa are not installable here): each step is BASELINE configs[0] in
         full - one 30 s stereo track, 4 pipeline chunks / 8 model windows and their features.

Beside the contract's keys the B200 line carries: ``stem_sdr_db`` (the timed 16-bit path's stems against the CPU oracle
on the cpu_baseline sample), ``fp32_value`` / ``bf16_value`` (the same device-timed step on the other arithmetic paths),
and ``configs`` = BASELINE configs[2] (one hour of mono audio, feature-only, host buffers in and out, per-kernel HBM
fractions; N = 1) and configs[4] (one 60-minute stereo mix chunk-sharded over the N ranks with halo recompute, merged
on the host, strong scaling, checked sample for sample against the single-GPU stitch).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 44100
TRACK_SECONDS = 240.0
METRIC = "x-realtime (audio s/s) separation+features"
UNIT = "audio-s/s"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm": float(p["hbm_gbs"]), "bf16_burst": float(p["bf16_tflops"]), "bf16_sustained": float(p["bf16_tflops_sustained"]),
                "src": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        return {"hbm": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled DURING the timed region (pynvml, 100 ms period)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()

    def run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = int(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = {
                nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
            }
            while not self._stop_evt.is_set():
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
                time.sleep(0.1)
        except Exception as exc:  # pragma: no cover
            self.reasons.add(f"sampler_error:{type(exc).__name__}")

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = int(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle restatement of the reference's CPU path
# ------------------------------------------------------------------------------------------------
def cpu_reference_pass(sample_seconds: float, n_fft: int, threads: int, state=None, want_stems: bool = False):
    """One bounded pass of the reference CPU path over `sample_seconds` of the workload: chunked
    MDX23 separation (backends.py:299-406 + enhanced_vocal_separator.py:366-458 restated) with the
    torch-CPU TFC-TDF net, ChunkFeatureBuilder features (features_cache.py:122-335 restated) and the
    vocal RMS / flatness series of PureVocalPauseDetector.  Returns (audio_seconds, wall_seconds)."""
    import torch

    from audio_cut_b200 import synth, unet_weights as uw
    from oracle import features as OF
    from oracle import mdx, pipeline, planner
    from oracle import unet as ounet

    torch.set_num_threads(threads)
    geo = uw.UNetGeometry()
    st = state if state is not None else uw.random_state(geo)
    net = ounet.build_net(st, geo.dim_f, geo.dim_t, geo.g)
    mg = mdx.MdxGeometry(n_fft=n_fft)
    audio = synth.synth_track(sample_seconds, seed=0, stereo=True)
    n = audio.shape[-1]
    plans = planner.chunk_schedule(n / float(SR))
    t0 = time.perf_counter()
    vocal, instr = pipeline.separate_track(audio, lambda ch: mdx.infer_chunk(ch, net, mg), sr=SR, plans=plans)
    mono = audio.mean(axis=0)
    cf = pipeline.ChunkFeatures(SR)
    for p in plans:
        cs, ce, _, _ = planner.sample_bounds(p, SR, n)
        cf.add_chunk(p, mono[cs:ce])
    out = cf.finalize()
    OF.onset_strength(out["bpm_wave"], SR, 512, aggregate=np.median)
    for fr, hop in ((1102, 441), (2048, 441), (2205, 882)):
        OF.rms(vocal, fr, hop)
    OF.spectral_flatness(vocal, 2048, 441)
    OF.rms(mono, 2048, 441)
    wall = time.perf_counter() - t0
    if want_stems:
        return n / float(SR), wall, (audio, plans, vocal, instr)
    return n / float(SR), wall


def bench_config(n_fft: int, world: int, chunks: int = 32, windows: int = 64) -> dict:
    """The workload description shared by both arms (the reference arm runs a bounded sample of it)."""
    return {"workload": "v2.2_mdd hot path, one 4-min 44.1 kHz stereo track per GPU per step (configs[1]; track-sharded at N>1 = configs[3])",
            "n_fft": n_fft, "hop": 1024, "dim_f": 3072, "dim_t": 256, "chunk_s": 10.0, "overlap_s": 2.5, "halo_s": 0.5,
            "chunks_per_track": chunks, "windows_per_track": windows, "weights": "random-init TFC-TDF (Kim_Vocal geometry, seed 1234)",
            "l2": "working set per step (>4 GB of activations, 85 MB track) exceeds the 126 MB L2; no explicit flush",
            "parallelism": f"track-sharded x{world}, no collective"}


REF_SAMPLE_S = 30.0  # BASELINE configs[0]: the reference's own CPU-runnable case, in full


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    from audio_cut_b200 import unet_weights as uw
    from oracle import mdx, planner

    st = uw.random_state(uw.UNetGeometry())
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_reference_pass(REF_SAMPLE_S, args.n_fft, threads, st)
    audio_s, wall = 0.0, 0.0
    for _ in range(args.steps):
        a, w = cpu_reference_pass(REF_SAMPLE_S, args.n_fft, threads, st)
        audio_s += a
        wall += w
    value = audio_s / wall
    plans = planner.chunk_schedule(REF_SAMPLE_S)
    mg = mdx.MdxGeometry(n_fft=args.n_fft)
    n_win = sum(mdx.n_windows(planner.sample_bounds(p, SR, int(REF_SAMPLE_S * SR))[1] - planner.sample_bounds(p, SR, int(REF_SAMPLE_S * SR))[0], mg)
                for p in plans)
    cfg = {"workload": "BASELINE configs[0] in full: one 30 s 44.1 kHz stereo track per step through the reference's CPU path "
                       "(chunked MDX23 separation + ChunkFeatureBuilder features + the vocal RMS / flatness series); a bounded "
                       "sample of the B200 arm's workload (4-min tracks, same chunk schedule and window geometry)",
           "sample_s": REF_SAMPLE_S, "chunks_per_step": len(plans), "windows_per_step": int(n_win),
           "n_fft": args.n_fft, "hop": 1024, "dim_f": 3072, "dim_t": 256, "chunk_s": 10.0, "overlap_s": 2.5, "halo_s": 0.5,
           "weights": "random-init TFC-TDF (Kim_Vocal geometry, seed 1234)", "arithmetic": "fp32 (torch-CPU), the B200 arm's default is fp16 operands / fp32 accumulation",
           "parallelism": f"{threads} host threads, rank 0 only"}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{REF_SAMPLE_S:.0f} s stereo (configs[0]: {len(plans)} chunks, {n_win} MDX windows + features) per step, "
                                   f"{args.steps} steps; oracle port: torch-CPU TFC-TDF net + numpy/scipy librosa restatement "
                                   "(onnxruntime / librosa absent from the image)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def device_step(net_backend, mix_dev, plans, bounds, series_out, dtype=None):
    """One hot-path pass with inputs resident in HBM; every series stays on the device."""
    import torch

    from audio_cut_b200 import ops
    from audio_cut_b200.features_cache import B200ChunkFeatureBuilder

    be = net_backend
    vocal, instr, _ = ops.separate_track(be.net, mix_dev, bounds, be.geom, align_hop=be.align_hop,
                                         output_is_vocal=True, dtype=be.dtype if dtype is None else dtype)
    mono = mix_dev.mean(dim=0)
    hop = 2205
    segs, off = [], 0
    for cs, ce, _, _ in bounds:
        segs.append((cs, ce - cs, off))
        off += 1 + (ce - cs) // hop
    feats = ops.stft_features(mono, segs, hop, SR, total_frames=off, want=("flatness", "onset_mean"))
    rsegs, roff = [], 0
    for cs, ce, _, _ in bounds:
        rsegs.append((cs, ce - cs, roff))
        roff += ops.frame_count(ce - cs, 4410, hop)
    rms_c = ops.frame_rms_segments(mono, rsegs, 4410, hop, total_frames=roff)  # all 32 chunks in one launch
    # BPM front end: onset envelope (median) at hop 512 over the effective-region concat
    bpm_wave = torch.cat([mono[es:ee] for _, _, es, ee in bounds])
    n_b = bpm_wave.numel()
    bpm_env = ops.stft_features(bpm_wave, [(0, n_b, 0)], 512, SR, total_frames=1 + n_b // 512, want=("onset_median",))
    # PureVocalPauseDetector / SeamlessSplitter series on the stems
    v_rms = [ops.frame_rms(vocal, fr, hp) for fr, hp in ((1102, 441), (2048, 441), (2205, 882))]
    v_flat = ops.stft_features(vocal, [(0, vocal.numel(), 0)], 441, SR, total_frames=1 + vocal.numel() // 441, want=("flatness",))
    m_rms = ops.frame_rms(mono, 2048, 441)
    series_out[:] = [vocal, instr, feats, rms_c, bpm_env, v_rms, v_flat, m_rms]


def time_device_steps(be, mix_dev, plans, bounds, dtype, steps, barrier):
    """ms per step of ``device_step`` on another arithmetic path (same network, same kernels' schedule)."""
    import torch

    keep = []
    device_step(be, mix_dev, plans, bounds, keep, dtype=dtype)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        device_step(be, mix_dev, plans, bounds, keep, dtype=dtype)
    e1.record()
    barrier()
    return e0.elapsed_time(e1) / steps


def sdr_db(ref, est):
    ref, est = np.asarray(ref, np.float64), np.asarray(est, np.float64)
    den = float(np.sum((ref - est) ** 2))
    return float("inf") if den == 0 else 10.0 * np.log10(float(np.sum(ref ** 2)) / den + 1e-300)


def features_1h_config(dev, peaks):
    """BASELINE configs[2]: every framewise series over one hour of mono audio, HOST buffers in and out."""
    import torch

    from audio_cut_b200 import featbench, ops, synth

    seconds = 3600.0
    base = synth.synth_track(240.0, seed=3, stereo=False)
    n = int(seconds * SR)
    host = torch.empty(n, dtype=torch.float32, pin_memory=True)
    host.numpy()[:] = np.tile(base, n // base.size + 1)[:n]
    nf441, nf512 = 1 + n // 441, 1 + n // 512
    outs_pin = {}

    def run():
        x = host.to(dev, non_blocking=True)
        res = {}
        for fr, hp in ((4410, 2205), (1102, 441), (2048, 441), (2205, 882)):
            res[f"rms_{fr}_{hp}"] = ops.frame_rms(x, fr, hp)
        res["zcr"] = ops.zero_crossing_rate(x, 2048, 441)
        res.update({"v_" + k: v for k, v in ops.stft_features(x, [(0, n, 0)], 441, SR, total_frames=nf441,
                                                                   want=("flatness", "centroid", "low_ratio")).items()})
        res.update({"b_" + k: v for k, v in ops.stft_features(x, [(0, n, 0)], 512, SR, total_frames=nf512,
                                                                   want=("onset_mean", "onset_median")).items()})
        f0, flag, vp = ops.pyin(x, SR, 441)
        res["f0"], res["voiced_prob"], res["voiced_flag"] = f0, vp, flag.to(torch.float32)
        mags, counts = ops.lpc_formants(x, SR, 441, 12)
        res["formant_mags"], res["formant_counts"] = mags.reshape(-1), counts.to(torch.float32)
        nbytes = 0
        for k, v in res.items():
            if k not in outs_pin:
                outs_pin[k] = torch.empty(v.numel(), dtype=torch.float32, pin_memory=True)
            outs_pin[k].copy_(v.reshape(-1), non_blocking=True)
            nbytes += v.numel() * 4
        torch.cuda.synchronize()
        return nbytes

    run()
    t0 = time.perf_counter()
    d2h = run()
    wall = time.perf_counter() - t0
    table = featbench.feature_only_bench(seconds, device=dev.index or 0, hbm_peak_gbs=peaks["hbm"])
    return {"workload": "BASELINE configs[2]: RMS x4 geometries, ZCR, STFT-2048 flatness/centroid/band ratio @441, BPM onset envelope "
                        "mean+median @512, pYIN (YIN + fp64 Viterbi), LPC formants over 1 h of synthetic mono audio",
            "e2e": {"value": seconds / wall, "unit": UNIT, "ms": wall * 1000.0, "h2d_bytes": int(n * 4), "d2h_bytes": int(d2h),
                    "call": "pinned host ndarray -> ops.* -> pinned host series"},
            "device_resident": {"value": table["x_realtime_all_series"], "unit": UNIT, "ms": table["total_ms"]},
            "kernels": table["kernels"], "hbm_peak_GBps": peaks["hbm"]}


def chunk_sharded_config(be, rank, world, dev, barrier, dist):
    """BASELINE configs[4]: ONE 60-minute stereo mix, pipeline chunks dealt to the ranks in contiguous blocks (halo
    recomputed locally, each rank uploads only the samples its chunks touch), shards merged on the host with the
    reference's accumulate / weight rule; strong scaling.  Rank 0 also runs the whole mix alone and compares."""
    import torch

    from audio_cut_b200 import ops, sharding, synth
    from audio_cut_b200.gpu_pipeline import chunk_schedule

    seconds = 3600.0
    base = synth.synth_track(240.0, seed=11, stereo=True)
    mix = np.ascontiguousarray(np.tile(base, (1, 15)))
    n = mix.shape[1]
    plans = chunk_schedule(n / float(SR))
    bounds = [p.sample_bounds(SR, n) for p in plans]
    kw = dict(align_hop=be.align_hop, output_is_vocal=True, dtype=be.dtype)
    lo_c, hi_c = sharding.shard_chunks(len(bounds), world)[rank]
    mine = list(bounds[lo_c:hi_c])
    lo, hi = sharding.shard_sample_range(mine)
    local_host = torch.from_numpy(np.ascontiguousarray(mix[:, lo:hi])).pin_memory()
    local_bounds = sharding.localize_bounds(mine, lo)

    def one_pass():
        local = local_host.to(dev, non_blocking=True)
        return ops.separate_track(be.net, local, local_bounds, be.geom, **kw)

    one_pass()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    v, i, w = one_pass()
    e1.record()
    torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1)
    shard = {"lo": lo, "vocal": v.cpu().numpy(), "instr": i.cpu().numpy(), "weight": w.cpu().numpy()}
    tag = os.environ.get("MASTER_PORT", "0")
    path = f"/dev/shm/acb200_{tag}_{rank}.npz"
    if world > 1 and rank != 0:
        np.savez(path + ".tmp.npz", **shard)
        os.replace(path + ".tmp.npz", path)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)  # timing only
        barrier()  # host-blocking: every rank's shard file is complete
    out = None
    if rank == 0:
        shards = [shard]
        for r in range(1, world):
            with np.load(f"/dev/shm/acb200_{tag}_{r}.npz") as z:
                shards.append({k: z[k] for k in z.files})
            os.remove(f"/dev/shm/acb200_{tag}_{r}.npz")
        mv, mi = sharding.merge_chunk_shards(n, shards)
        e2e_s = time.perf_counter() - t0
        # the single-GPU stitch of the same mix, for the identity check and the strong-scaling denominator
        full = torch.from_numpy(mix).to(dev)
        ops.separate_track(be.net, full, bounds, be.geom, **kw)
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        fv, fi, _ = ops.separate_track(be.net, full, bounds, be.geom, **kw)
        s1.record()
        torch.cuda.synchronize()
        t1_ms = s0.elapsed_time(s1)
        same = bool(np.array_equal(mv, fv.cpu().numpy()) and np.array_equal(mi, fi.cpu().numpy()))
        tn_ms = float(t[0])
        out = {"workload": "BASELINE configs[4]: one 60-min 44.1 kHz stereo mix, 480 chunks / 960 windows, chunk-sharded with halo recompute, "
                           "host merge (no collective on the data path)", "scaling": "strong", "n_gpus": world,
               "value": seconds / (tn_ms / 1000.0), "unit": UNIT, "ms": tn_ms, "timing": "CUDA events, max over ranks (upload of the rank's samples + separation)",
               "e2e_value": seconds / e2e_s, "e2e_ms": e2e_s * 1000.0, "e2e_includes": "D2H of every shard, /dev/shm hand-over, host merge on rank 0",
               "single_gpu_ms": t1_ms, "strong_scaling_efficiency": t1_ms / (world * tn_ms),
               "merged_equals_single_gpu_stitch": same, "chunks_per_rank": [b - a for a, b in sharding.shard_chunks(len(bounds), world)]}
    if world > 1:
        dist.barrier()
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist

    from audio_cut_b200 import _lib, synth, unet_weights as uw
    from audio_cut_b200.backends import B200Mdx23Backend
    from audio_cut_b200.gpu_pipeline import PipelineConfig, chunk_schedule
    from audio_cut_b200.separator import B200VocalSeparator

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the B200 arm has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its version banner to stdout when the first communicator is created; rank 0 must print
        # exactly ONE JSON line there, so file descriptor 1 points at stderr until the communicator exists
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    dev = torch.device("cuda", local)

    geo = uw.UNetGeometry()
    state = uw.random_state(geo)
    be = B200Mdx23Backend(weights=state, device=f"cuda:{local}", precision=args.precision, n_fft=args.n_fft, output_type="vocal")
    be.load_model()
    # independent tracks per rank (track sharding): different seed per rank, same length
    audio = synth.synth_track(TRACK_SECONDS, seed=rank, stereo=True)
    n = audio.shape[-1]
    plans = chunk_schedule(n / float(SR))
    bounds = [p.sample_bounds(SR, n) for p in plans]
    mix_dev = torch.from_numpy(audio).to(dev)
    import ctypes

    from audio_cut_b200 import ops
    n_windows = int(_lib.load().ac_track_window_count(ops.make_chunk_descs(bounds), len(bounds),
                                                      ctypes.byref(_lib.TrackParams(be.geom, be.align_hop, 2, 1, be.dtype, 0, 0))))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    keep = []
    for _ in range(args.warmup):
        device_step(be, mix_dev, plans, bounds, keep)
    barrier()
    lib = _lib.load()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    _lib.profile_begin()
    launches0 = lib.ac_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        device_step(be, mix_dev, plans, bounds, keep)
    e1.record()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = int(lib.ac_launch_count() - launches0)
    kstats = _lib.profile_collect()
    clocks = sampler.stop() if sampler else None

    # ---- e2e through the reference-facing call, host buffers in / out
    # The step's input lives in page-locked host memory (the contract's "host->device copy of that step's inputs from
    # pinned host memory"); the drop-in recognises that and skips its own pageable->pinned staging.  The same loop
    # with an ordinary pageable ndarray is timed as well and reported beside it.
    sep = B200VocalSeparator(SR, backend=be, pipeline_config=PipelineConfig())
    audio_pin_t = torch.empty(audio.shape, dtype=torch.float32, pin_memory=True)
    audio_pin = audio_pin_t.numpy()
    audio_pin[...] = audio

    def e2e_loop(host_audio):
        for _ in range(min(args.warmup, 2)):
            r = sep.separate_for_detection(host_audio)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            r = sep.separate_for_detection(host_audio)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        barrier()
        return r, dt

    _, e2e_pageable_s = e2e_loop(audio)
    res, e2e_s = e2e_loop(audio_pin)
    fc = res.feature_cache
    d2h = res.vocal_track.nbytes + (res.instrumental_track.nbytes if res.instrumental_track is not None else 0) + \
        4 * (fc.rms_series.size + fc.spectral_flatness.size + fc.onset_envelope.size) + 4 * (1 + (n + 1323000) // 512) + 4 * (1 + n // 882)

    # ---- the other arithmetic paths (N = 1), configs[2] (N = 1) and configs[4] (every N)
    other = {}
    if world == 1 and not args.quick:
        for name, dt in (("fp16", _lib.AC_F16), ("bf16", _lib.AC_BF16), ("fp32", _lib.AC_F32)):
            if name != args.precision:
                other[name] = time_device_steps(be, mix_dev, plans, bounds, dt, 2, barrier)
    extra_cfgs = {}
    if not args.quick:
        if world == 1:
            extra_cfgs["features_1h"] = features_1h_config(dev, _peaks())
        cs = chunk_sharded_config(be, rank, world, dev, barrier, dist)
        if cs is not None:
            extra_cfgs["chunk_sharded_60min"] = cs

    t = torch.tensor([dev_ms, e2e_s * 1000.0, e2e_pageable_s * 1000.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, e2e_pageable_ms = float(t[0]), float(t[1]), float(t[2])
    if rank == 0:
        peaks = _peaks()
        audio_total = TRACK_SECONDS * world * args.steps
        value = audio_total / (dev_ms / 1000.0)
        e2e_value = audio_total / (e2e_ms / 1000.0)
        kstats.sort(key=lambda k: -k["total_ms"])
        for k in kstats:
            k["share"] = k["total_ms"] / max(1e-9, sum(x["total_ms"] for x in kstats))
            k["avg_ms"] = k["total_ms"] / max(1, k["launches"])
            k["tflops"] = k["flops"] / (k["total_ms"] * 1e9) if k["total_ms"] > 0 else 0.0
            k["gbs"] = k["bytes"] / (k["total_ms"] * 1e6) if k["total_ms"] > 0 else 0.0
        top = kstats[0]
        tensor_bound = top["flops"] > 0
        traffic, traffic_src = None, None  # DRAM bytes per launch of the dominant kernel class, from the committed ncu --set full capture
        try:
            for name in ("r02_traffic.json", "r01_traffic.json"):
                p = os.path.join(ROOT, "profiles", name)
                if os.path.exists(p):
                    with open(p) as f:
                        traffic = float(json.load(f)[top["name"]]["dram_bytes_per_launch"])
                    traffic_src = f"profiles/{name} (ncu --set full capture of the same kernels, dram__bytes_read.sum + dram__bytes_write.sum per launch)"
                    break
        except Exception:
            traffic = None
        if tensor_bound:
            roof = {"bound": "tensor", "kernel": top["name"], "achieved": top["tflops"], "peak": peaks["bf16_sustained"],
                    "unit": "TFLOP/s", "frac": top["tflops"] / peaks["bf16_sustained"], "traffic": traffic,
                    "traffic_source": traffic_src,
                    "peak_source": peaks["src"] + ", sustained bf16 = f16 tensor-core rate (kernel timed inside a long step)",
                    "avg_launch_ms": top["avg_ms"], "share_of_step": top["share"]}
        else:
            roof = {"bound": "hbm", "kernel": top["name"], "achieved": top["gbs"], "peak": peaks["hbm"], "unit": "GB/s",
                    "frac": top["gbs"] / peaks["hbm"], "traffic": traffic, "peak_source": peaks["src"],
                    "avg_launch_ms": top["avg_ms"], "share_of_step": top["share"]}
        cpu_threads = os.cpu_count() or 1
        cpu, stem_sdr = None, None
        if world == 1 and not args.no_cpu_baseline:
            a, w, (s_audio, s_plans, s_vocal, s_instr) = cpu_reference_pass(REF_SAMPLE_S, args.n_fft, cpu_threads, state, want_stems=True)
            cpu = {"value": a / w, "unit": UNIT, "cores": cpu_threads, "kind": "port",
                   "sample": f"{REF_SAMPLE_S:.0f} s stereo (BASELINE configs[0]: {len(s_plans)} chunks, 8 MDX windows + features), one pass; oracle "
                             "port: torch-CPU TFC-TDF net + numpy/scipy librosa restatement (onnxruntime / librosa absent from the image)"}
            # the timed path's stems against that CPU pass (the oracle is the checker here, never the thing measured)
            s_bounds = [p.sample_bounds(SR, s_audio.shape[-1]) for p in chunk_schedule(s_audio.shape[-1] / float(SR))]
            gv, gi, _ = ops.separate_track(be.net, torch.from_numpy(s_audio).to(dev), s_bounds, be.geom, align_hop=be.align_hop,
                                           output_is_vocal=True, dtype=be.dtype)
            stem_sdr = {"vocal": sdr_db(s_vocal, gv.cpu().numpy()), "instrumental": sdr_db(s_instr, gi.cpu().numpy()),
                        "path": args.precision, "vs": f"CPU oracle, {REF_SAMPLE_S:.0f} s stereo, n_fft {args.n_fft}",
                        "gate": "north_star: >= 40 dB for the 16-bit path, >= 60 dB for fp32"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp16": "f16", "bf16": "bf16", "fp32": "f32"}[args.precision], "data": "synthetic",
            "config": bench_config(args.n_fft, world, len(plans), n_windows),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(audio.nbytes), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms / args.steps,
                    "call": "B200VocalSeparator.separate_for_detection(page-locked host ndarray) -> host ndarrays",
                    "ms_per_step_pageable_input": e2e_pageable_ms / args.steps},
            "gpu_launches": launches,
            "roofline": roof,
            "kernels": [{k2: (round(v, 4) if isinstance(v, float) else v) for k2, v in k.items()} for k in kstats],
            "unet_tflops_overall": sum(k["flops"] for k in kstats) / (sum(k["total_ms"] for k in kstats if k["flops"] > 0) * 1e9 + 1e-9),
            "cpu_baseline": cpu,
            "stem_sdr_db": stem_sdr,
            "configs": extra_cfgs,
        }
        for name, ms in other.items():
            line[f"{name}_value"] = TRACK_SECONDS / (ms / 1000.0)
            line[f"{name}_ms_per_step"] = ms
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16", "fp32"])
    ap.add_argument("--n-fft", dest="n_fft", type=int, default=7680)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="headline workload only: skip the other precisions and BASELINE configs[2] / configs[4]")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
