python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "stft or istft or feature" 2>&1 | tail -3
for b in 8 16 32 64; do python scripts/dev_unet_tc_once.py $b fp16 3 2>&1 | tail -1; done
