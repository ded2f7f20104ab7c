set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "bench rc=$?"
python scripts/dev_unet_tc_once.py 4 > gpurun_out/plain_unet.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tc_ -s 57 -c 8 -o gpurun_out/prof_unet_full python scripts/dev_unet_tc_once.py 4 > gpurun_out/ncu_unet_full.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/plain_unet.log
