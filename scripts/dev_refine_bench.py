"""Dev helper: finalize_cut_points on a 4-minute track (120 candidates, default parameters) - the GPU drop-in
(stems resident in HBM) against the CPU restatement of the reference (oracle/cuts.py; it omits the reference's
O(N) pure-Python `next_quiet` loop, so the reference itself is slower still).  Writes profiles/r01_refine.json."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from audio_cut_b200 import refine as R, synth, ops
from oracle import cuts

SR = 44100
mix = synth.synth_track(240.0, seed=11).mean(axis=0).astype(np.float32)
vocal = (0.6 * mix * (np.sin(2 * np.pi * 0.2 * np.arange(mix.size) / SR) > 0)).astype(np.float32)
rng = np.random.default_rng(5)
pts = [(float(a), float(b)) for a, b in zip(rng.uniform(0.3, 239.7, 120), rng.uniform(0, 1, 120))]
kw = dict(min_gap_s=1.0, topk_per_10s=6, floor_db=-50.0)
d_mix, d_vocal = torch.from_numpy(mix).cuda(), torch.from_numpy(vocal).cuda()
ctx = R.CutContext(sr=SR, mix_wave=d_mix, vocal_wave=d_vocal)
cps = [R.CutPoint(a, b) for a, b in pts]
for _ in range(3):
    res = R.finalize_cut_points(ctx, cps, **kw)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    res = R.finalize_cut_points(ctx, cps, **kw)
t_gpu = (time.perf_counter() - t0) / 20
# kernel alone
pruned = R.nms_min_gap(cps, 1.0, max_per_window=6)
times = [p.t for p in pruned]
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
t_in = None
e0.record()
for _ in range(20):
    ops.refine_cut_points(d_mix, d_vocal, SR, times, zero_cross_half=353, search=6615, win=441, guard_db=2.0, floor_db=-50.0)
e1.record(); torch.cuda.synchronize()
t_call = e0.elapsed_time(e1) / 20
t0 = time.perf_counter()
bounds, ktimes = cuts.finalize_cut_points(mix, vocal, SR, pts, **kw)
t_cpu = time.perf_counter() - t0
assert bounds == res.sample_boundaries
out = {"track_s": 240.0, "candidates": len(pts), "pruned": len(pruned), "kept": len(ktimes),
       "gpu_finalize_cut_points_ms": t_gpu * 1e3, "gpu_refine_call_ms_incl_d2h": t_call,
       "cpu_oracle_finalize_cut_points_ms": t_cpu * 1e3, "boundaries_equal": True}
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/r01_refine.json", "w"), indent=1)
