"""Dev helper: feature-only path (BASELINE configs[2]) on 1 h of synthetic mono audio: per-kernel time and GB/s."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_cut_b200.featbench import feature_only_bench
print(json.dumps(feature_only_bench(float(sys.argv[1]) if len(sys.argv) > 1 else 3600.0), indent=1))
