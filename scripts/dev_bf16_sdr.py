"""Stem SDR of the bf16 tcgen05 path vs the CPU oracle at full Kim_Vocal geometry (BASELINE configs[0]).

    python scripts/dev_bf16_sdr.py [seconds] [n_fft]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import sdr_db  # noqa: E402

from audio_cut_b200 import _lib, ops, synth, unet_weights as uw  # noqa: E402
from oracle import mdx, pipeline, planner  # noqa: E402
from oracle import unet as ounet  # noqa: E402

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
n_fft = int(sys.argv[2]) if len(sys.argv) > 2 else 7680
sr = 44100
geo = uw.UNetGeometry()
st = uw.random_state(geo, seed=1234)
net = ops.UNet(st, geo)
ref_net = ounet.build_net(st, geo.dim_f, geo.dim_t, geo.g)
mg = mdx.MdxGeometry(n_fft, 1024, 3072, 256)
audio = synth.synth_track(seconds, sr=sr, seed=0, stereo=True)
total = audio.shape[-1]
plans = planner.chunk_schedule(total / float(sr), 10.0, 2.5, 0.5)
bounds = [planner.sample_bounds(p, sr, total) for p in plans]
t0 = time.time()
ref_v, ref_i = pipeline.separate_track(audio, lambda ch: mdx.infer_chunk(ch, ref_net, mg, align_hop=4096), sr=sr, plans=plans)
t_or = time.time() - t0
mix = torch.from_numpy(audio).cuda()
geom = ops.mdx_geom(n_fft, 1024, 3072, 256)
out = {"seconds": seconds, "n_fft": n_fft, "oracle_s": t_or, "chunks": len(bounds)}
for name, dt in (("fp32", _lib.AC_F32), ("bf16", _lib.AC_BF16), ("fp16", _lib.AC_F16)):
    v, i, w = ops.separate_track(net, mix, bounds, geom, dtype=dt)
    v, i = v.cpu().numpy(), i.cpu().numpy()
    out[name] = {"vocal_sdr": float(sdr_db(ref_v, v)), "instr_sdr": float(sdr_db(ref_i, i)),
                 "vocal_rms": float(np.sqrt(np.mean(ref_v.astype(np.float64) ** 2))),
                 "instr_rms": float(np.sqrt(np.mean(ref_i.astype(np.float64) ** 2)))}
# raw network SDR on one window of real spectrogram
x = torch.randn(1, 4, 3072, 256, generator=torch.Generator().manual_seed(0)) * 3.0
with torch.no_grad():
    ref = ref_net(x)
got = ops.tfc_to_onnx(net.forward(ops.onnx_to_tfc(x.cuda()).bfloat16())).float().cpu()
out["net_bf16_sdr_randn"] = float(sdr_db(ref.numpy(), got.numpy()))
got = ops.tfc_to_onnx(net.forward(ops.onnx_to_tfc(x.cuda()).half())).float().cpu()
out["net_fp16_sdr_randn"] = float(sdr_db(ref.numpy(), got.numpy()))
out["aborted"] = int(_lib.load().ac_debug_tc_aborted())
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "bf16_sdr.json"), "w") as f:
    json.dump(out, f)
