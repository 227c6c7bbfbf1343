set -x
for p in 1 0 1 0; do
  timeout 300 python tools/profile_step.py --batch 64 --set gemm_pair=$p 2>&1 | grep -E "gpu_ms|conv_tc_tf32x3" 
done
for p in 1 0; do
  KKX_LIB=kokorox_b200/lib/libkkx_timing.so KKX_ARB_TIMING=1 KKX_ARB_TIMING_KS=99 timeout 300 python tools/profile_step.py --batch 64 --runs 2 --set gemm_pair=$p 2>&1 | grep "tf32x3 timing" | tail -1
done
