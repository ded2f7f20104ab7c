# Round-end measurement pass (run under gpurun): tests, both bench arms, ncu launch lists, one full-set capture.
set -x
R=${1:-r01}
python -m pytest tests -m gpu -x -q > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${R}_pytest_gpu.log
tail -2 gpurun_out/${R}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${R}_smoke.log 2>&1; tail -1 gpurun_out/${R}_smoke.log
python bench.py > gpurun_out/${R}_bench_final.json 2> gpurun_out/${R}_bench_final.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_bench_reference.err; echo "ref rc=$?"
python scripts/dev_feature_bench.py 3600 > gpurun_out/${R}_features_1h.json 2>/dev/null; echo "features rc=$?"
python scripts/dev_refine_bench.py > /dev/null 2>&1; cp gpurun_out/r01_refine.json gpurun_out/${R}_refine.json 2>/dev/null
# launch list of the bench command itself (times under ncu are cold-cache and serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/${R}_launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${R}_ncu_bench.log 2>&1
python scripts/ncu_summary.py list gpurun_out/${R}_launches_bench.csv > gpurun_out/${R}_bench_launches.md
# U-Net forward launch list (16 windows)
python scripts/dev_unet_tc_once.py 16 > gpurun_out/${R}_plain_unet16.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches_unet.csv \
    python scripts/dev_unet_tc_once.py 16 > /dev/null 2>&1
python scripts/launch_table.py gpurun_out/${R}_launches_unet.csv > gpurun_out/${R}_unet_launches.txt
# full-set capture of every tensor-core kernel of one forward (second forward), report kept on the box, summary exported
# (55 tensor-core launches per forward since the level-0 conv chains are one launch each)
ncu --set full --clock-control none --import-source on -k regex:"tc_conv3x3|tc_tdf|tc_resample" -s 55 -c 55 -o /tmp/prof_unet \
    python scripts/dev_unet_tc_once.py 16 > gpurun_out/${R}_ncu_unet_full.log 2>&1
python scripts/ncu_summary.py rep /tmp/prof_unet.ncu-rep > gpurun_out/${R}_unet_ncu_full.md
python scripts/make_traffic_json.py gpurun_out/${R}_unet_ncu_full.md gpurun_out/${R}_traffic.json
# full-set capture of the MDX STFT / fused iSTFT (three-pass FFT) and the feature kernels
python scripts/dev_hbm_once.py > gpurun_out/${R}_hbm_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"stft_mdx|istft_mdx|frame_reduce|stft_feat|onset_flux" -c 40 -o /tmp/prof_hbm \
    python scripts/dev_hbm_once.py > gpurun_out/${R}_ncu_hbm.log 2>&1
python scripts/ncu_hbm_summary.py /tmp/prof_hbm.ncu-rep > gpurun_out/${R}_hbm_ncu_full.md 2>/dev/null || python scripts/ncu_summary.py rep /tmp/prof_hbm.ncu-rep > gpurun_out/${R}_hbm_ncu_full.md
rm -f gpurun_out/${R}_launches_bench.csv
cat gpurun_out/${R}_bench_final.json | head -c 600; echo; cat gpurun_out/${R}_plain_unet16.log
