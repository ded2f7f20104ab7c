# one gpurun call: (1) ncu --set full of the third (timed) 16-window fp16 U-Net forward, (2) ncu launch list of a short bench
python scripts/dev_unet_tc_once.py 16 fp16 > gpurun_out/unet_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'tc_|first_conv|final_conv|tdf_small' -s 132 -c 66 \
  -o gpurun_out/r02_unet -f python scripts/dev_unet_tc_once.py 16 fp16 > gpurun_out/unet_ncu.log 2>&1
tail -2 gpurun_out/unet_plain.log
python bench.py --steps 1 --warmup 1 --quick --no-cpu-baseline > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_bench_launches.csv \
  python bench.py --steps 1 --warmup 1 --quick --no-cpu-baseline > gpurun_out/bench_ncu.json 2> gpurun_out/bench_ncu.err
ls -la gpurun_out/r02_unet.ncu-rep gpurun_out/r02_bench_launches.csv
