set -x
KKX_LIB=kokorox_b200/lib/libkkx_exp.so KKX_ARB_PB=1 timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "arb" 2>&1 | tail -2
for p in 1 0 1 0; do
  KKX_LIB=kokorox_b200/lib/libkkx_exp.so KKX_ARB_PB=$p KKX_PROFILE_DETAIL=1 timeout 300 python tools/profile_step.py --batch 64 2>&1 | grep -E "gpu_ms|arb_conv\[c128 k(3|7|11) d1 conv1 m5746720" | cut -c1-100
  KKX_LIB=kokorox_b200/lib/libkkx_exp.so KKX_ARB_PB=$p timeout 300 python tools/profile_step.py --batch 64 2>&1 | grep -E "^arb_conv "
done
