set -x
# 1. memcheck over every kernel family at small sizes
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_run.py > gpurun_out/r2_memcheck_v13.log 2>&1; echo "memcheck rc=$?"
tail -5 gpurun_out/r2_memcheck_v13.log
# 2. what bounds the fused res-block convs: skip loads / stores / transform / MMAs (experiment build, wrong results)
for d in 0 1 2 4 8 3 7 15; do
  KKX_LIB=kokorox_b200/lib/libkkx_exp.so KKX_ARB_DBG=$d KKX_PROFILE_DETAIL=1 timeout 300 python tools/profile_step.py --batch 64 2>&1 | grep -E "arb_conv\[c128 k(3|7|11) d1 conv[12] m5746720|arb_conv\[c256 k(3|7|11) d1 conv[12] m957780|launches_per_run" > gpurun_out/r2_arb_dbg${d}_v13.txt
  echo "dbg=$d"; cat gpurun_out/r2_arb_dbg${d}_v13.txt
done
# 3. role counters, k = 3 only, then k = 11
for k in 3 7 11; do
  KKX_LIB=kokorox_b200/lib/libkkx_timing.so KKX_ARB_TIMING=1 KKX_ARB_TIMING_KS=$k timeout 300 python tools/profile_step.py --batch 64 --runs 2 2>&1 | grep "arb timing" | tail -3 > gpurun_out/r2_arb_roles_k${k}_v13.txt
  cat gpurun_out/r2_arb_roles_k${k}_v13.txt
done
