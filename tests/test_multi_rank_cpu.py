"""N > 1 host logic on CPU: two gloo ranks chunk-shard one track and the host stitch equals the
single-process result.  The per-chunk 'separation' is a deterministic numpy
stand-in - the CUDA path is covered by the -m gpu tests; this checks partitioning and seams."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pipeline, planner


def _fake_infer(chunk):
    v = (0.7 * chunk + 0.01 * np.sin(np.arange(chunk.shape[-1], dtype=np.float32) * 0.01)).astype(np.float32)
    return v, (chunk - v).astype(np.float32)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, audio, sr, out_q):
    from audio_cut_b200 import sharding
    from audio_cut_b200.gpu_pipeline import chunk_schedule

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = len(audio)
    plans = chunk_schedule(n / float(sr), chunk_s=2.0, overlap_s=0.5, halo_s=0.1)
    bounds = [p.sample_bounds(sr, n) for p in plans]
    lo_c, hi_c = sharding.shard_chunks(len(plans), world)[rank]
    mine = bounds[lo_c:hi_c]
    lo, hi = sharding.shard_sample_range(mine)
    local = audio[lo:hi]
    vacc = np.zeros(hi - lo, np.float32)
    iacc = np.zeros(hi - lo, np.float32)
    wacc = np.zeros(hi - lo, np.float32)
    for cs, ce, es, ee in sharding.localize_bounds(mine, lo):
        v, i = _fake_infer(local[cs:ce])
        vacc[es:ee] += v[es - cs : ee - cs]
        iacc[es:ee] += i[es - cs : ee - cs]
        wacc[es:ee] += 1.0
    wn = np.where(wacc == 0, 1.0, wacc).astype(np.float32)
    shard = {"lo": lo, "vocal": vacc / wn, "instr": iacc / wn, "weight": wacc}
    gathered = [None] * world
    dist.all_gather_object(gathered, shard)
    dist.barrier()
    if rank == 0:
        out_q.put(sharding.merge_chunk_shards(n, gathered))
    dist.destroy_process_group()


def test_two_rank_chunk_sharding_matches_single_process():
    sr = 4000
    rng = np.random.default_rng(3)
    audio = (0.3 * rng.standard_normal(int(19.3 * sr))).astype(np.float32)
    plans = planner.chunk_schedule(len(audio) / float(sr), 2.0, 0.5, 0.1)
    ref_v, ref_i = pipeline.separate_track(audio, _fake_infer, sr=sr, plans=plans)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, audio, sr, q)) for r in range(2)]
    for p in procs:
        p.start()
    v, i = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    np.testing.assert_array_equal(v, ref_v)
    np.testing.assert_array_equal(i, ref_i)


def test_track_sharding_is_balanced():
    from audio_cut_b200 import sharding

    durs = [240.0] * 64
    for world in (1, 2, 4, 8):
        parts = sharding.shard_tracks(durs, world)
        assert sorted(sum(parts, [])) == list(range(64))
        assert {len(p) for p in parts} == {64 // world}
    parts = sharding.shard_tracks([300, 10, 200, 190, 20], 2)
    loads = [sum([300, 10, 200, 190, 20][i] for i in p) for p in parts]
    assert abs(loads[0] - loads[1]) <= 60
    assert sharding.shard_chunks(480, 8) == [(60 * r, 60 * (r + 1)) for r in range(8)]
    assert sharding.shard_chunks(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
