for sb in "16,16,16" "1,16,16" "2,16,16" "4,16,16" "1,2,16" "1,2,4" "1,2,8" "2,4,8" "1,1,2" "1,4,16" "8,16,16"; do
  AC_UNET_SB=$sb python scripts/dev_unet_tc_once.py 16 fp16 5 2>&1 | tail -1
done
python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "unet or track" 2>&1 | tail -3
