import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audio_cut_b200 import synth, unet_weights as uw
from audio_cut_b200.backends import B200Mdx23Backend
from audio_cut_b200.separator import B200VocalSeparator
be = B200Mdx23Backend(weights=uw.random_state(uw.UNetGeometry()), precision="bf16", output_type="vocal"); be.load_model()
sep = B200VocalSeparator(44100, backend=be)
audio = synth.synth_track(240.0, seed=0, stereo=True)
for _ in range(2): sep.separate_for_detection(audio)
t0=time.perf_counter(); res = sep.separate_for_detection(audio); print("e2e s", time.perf_counter()-t0, res.gpu_meta["gpu_pipeline_h2d_ms"], res.gpu_meta["gpu_pipeline_compute_ms"], res.gpu_meta["gpu_pipeline_dtoh_ms"])
pr = cProfile.Profile(); pr.enable(); sep.separate_for_detection(audio); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
