# one gpurun call: plain run, then ncu --set full on the HBM-class kernels (second pass of the script)
python scripts/dev_hbm_once.py > gpurun_out/hbm_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on \
  -k regex:'stft_mdx_kernel|istft_mdx_kernel|stft_feat|onset_flux|frame_reduce|yin_probs|lpc_formant' -s 16 -c 16 \
  -o gpurun_out/r02_hbm -f python scripts/dev_hbm_once.py > gpurun_out/hbm_ncu.log 2>&1
ls -la gpurun_out/r02_hbm.ncu-rep
