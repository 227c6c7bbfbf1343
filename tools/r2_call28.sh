set -x
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "bit_identical or split_fp16" 2>&1 | tail -2
timeout 300 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "pair_gemm or benched" 2>&1 | tail -2
for f in 0 1 0 1; do
  KKX_LIB=kokorox_b200/lib/libkkx_exp.so KKX_G2_FIN4=$f KKX_PROFILE_DETAIL=1 timeout 300 python tools/profile_step.py --batch 64 2>&1 | grep -E "gpu_ms|tf32x3\[ci768 co2304|tf32x3\[ci768 co2048|tf32x3\[ci2048|tf32x3\[ci768 co768" | cut -c1-100
  KKX_LIB=kokorox_b200/lib/libkkx_exp.so KKX_G2_FIN4=$f timeout 300 python tools/profile_step.py --batch 64 2>&1 | grep -E "conv_tc_tf32x3"
done
