python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "stft or istft or feature or track_small or full_geometry_30s_stereo or pyin or lpc" 2>&1 | tail -3
python scripts/dev_hbm_once.py > gpurun_out/hbm_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'stft_mdx_kernel|istft_mdx_kernel' -s 2 -c 2 \
  -o gpurun_out/r02_stft -f python scripts/dev_hbm_once.py > gpurun_out/stft_ncu.log 2>&1
python bench.py --steps 5 --warmup 3 --quick --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_q.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'], d['clocks'])
for k in d['kernels']: print(k['name'], k['launches'], round(k['total_ms'],2), round(k['avg_ms'],4), round(k['gbs']), round(k['tflops']))
"
