set -x
timeout 300 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "pair_gemm or phase_fused" 2>&1 | tail -8
for p in 1 0 1 0; do
  KKX_PROFILE_DETAIL=1 timeout 300 python tools/profile_step.py --batch 64 --set conv_pair=$p 2>&1 | grep -E "gpu_ms|conv_tc\[" | head -16
  timeout 300 python tools/profile_step.py --batch 64 --set conv_pair=$p 2>&1 | grep -E "^conv_tc "
done
