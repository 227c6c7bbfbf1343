set -x
timeout 300 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "pair_gemm" 2>&1 | tail -8
for p in 1 0 1 0; do
  KKX_PROFILE_DETAIL=1 timeout 300 python tools/profile_step.py --batch 64 --set conv_pair=$p 2>&1 | grep -E "gpu_ms|conv_tc\[ci(1024|1090|514|512|264)" | head -14
done
