"""Multi-GPU partitioning of the hot path: independent units, no collective.

* track sharding (BASELINE configs[3]): whole tracks are dealt to ranks, longest first;
* chunk sharding (configs[4]): one long track, each rank takes a contiguous block of pipeline
  chunks, uploads only the samples those chunks touch, and returns its stems for that sample range
  together with the overlap weights; the host stitches the seams with the reference's own rule
  (sum of accumulators / sum of weights, enhanced_vocal_separator.py:456-458) so the result is the
  same as the single-GPU one.  Halo frames are recomputed locally - nothing is exchanged between GPUs.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np


def shard_tracks(durations: Sequence[float], world: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of track indices to ranks."""
    loads = [0.0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for idx in sorted(range(len(durations)), key=lambda i: -durations[i]):
        r = min(range(world), key=lambda k: (loads[k], k))
        out[r].append(idx)
        loads[r] += durations[idx]
    return out


def shard_chunks(n_chunks: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous [lo, hi) chunk ranges, sizes differing by at most one."""
    base, extra = divmod(n_chunks, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def shard_sample_range(bounds: Sequence[Tuple[int, int, int, int]]) -> Tuple[int, int]:
    """Track samples [lo, hi) a block of chunks reads (chunk extents) - what the rank uploads."""
    if not bounds:
        return 0, 0
    return min(b[0] for b in bounds), max(b[1] for b in bounds)


def localize_bounds(bounds: Sequence[Tuple[int, int, int, int]], lo: int) -> List[Tuple[int, int, int, int]]:
    return [(cs - lo, ce - lo, es - lo, ee - lo) for cs, ce, es, ee in bounds]


def merge_chunk_shards(total_samples: int, shards: Sequence[Dict[str, np.ndarray]]):
    """shards: {"lo", "vocal", "instr", "weight"} per rank (stems already divided by max(weight,1)).

    accum = stem * weight is exact in float32 for weights 1, 2, 4 (the default schedule never
    exceeds 2), so summing accumulators and weights across ranks reproduces the single-GPU stitch.
    """
    vacc = np.zeros(total_samples, np.float32)
    iacc = np.zeros(total_samples, np.float32)
    wacc = np.zeros(total_samples, np.float32)
    for sh in sorted(shards, key=lambda s: int(s["lo"])):
        lo = int(sh["lo"])
        w = np.asarray(sh["weight"], np.float32)
        hi = lo + w.shape[0]
        vacc[lo:hi] += np.asarray(sh["vocal"], np.float32) * w
        iacc[lo:hi] += np.asarray(sh["instr"], np.float32) * w
        wacc[lo:hi] += w
    wacc[wacc == 0.0] = 1.0
    return vacc / wacc, iacc / wacc


def separate_chunk_shard(net, geom, mix: np.ndarray, bounds: Sequence[Tuple[int, int, int, int]], rank: int, world: int, *,
                         device: int = 0, **separate_kwargs) -> Dict[str, np.ndarray]:
    """What ONE rank does for a chunk-sharded track (BASELINE configs[4]): take its contiguous block of
    pipeline chunks, upload only the samples they touch, separate them on this rank's GPU with the same
    ``ac_separate_track`` entry point, and hand back the shard ``merge_chunk_shards`` expects.

    mix: host array [n_ch, N]; bounds: per-chunk (chunk_start, chunk_end, eff_start, eff_end) in samples for
    the WHOLE track (``ChunkPlan.sample_bounds``).  No collective: the caller gathers the dicts on the host.
    """
    import torch

    from . import ops

    lo_c, hi_c = shard_chunks(len(bounds), world)[rank]
    mine = list(bounds[lo_c:hi_c])
    lo, hi = shard_sample_range(mine)
    if hi <= lo:
        z = np.zeros(0, np.float32)
        return {"lo": 0, "vocal": z, "instr": z, "weight": z}
    local = torch.from_numpy(np.ascontiguousarray(mix[:, lo:hi], dtype=np.float32)).cuda(device)
    vocal, instr, weight = ops.separate_track(net, local, localize_bounds(mine, lo), geom, **separate_kwargs)
    torch.cuda.synchronize(local.device)
    return {"lo": lo, "vocal": vocal.cpu().numpy(), "instr": instr.cpu().numpy(), "weight": weight.cpu().numpy()}
