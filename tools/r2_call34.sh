set -x
KKX_LIB=kokorox_b200/lib/libkkx_exp2.so timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "arb" 2>&1 | tail -2
for l in exp2 exp exp2 exp; do
  KKX_LIB=kokorox_b200/lib/libkkx_$l.so KKX_PROFILE_DETAIL=1 timeout 300 python tools/profile_step.py --batch 64 2>&1 | grep -E "gpu_ms|arb_conv\[c128 k(3|7|11) d1 conv2 m5746720" | cut -c1-100
  KKX_LIB=kokorox_b200/lib/libkkx_$l.so timeout 300 python tools/profile_step.py --batch 64 2>&1 | grep -E "^arb_conv "
done
