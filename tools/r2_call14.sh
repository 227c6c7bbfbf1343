set -x
timeout 900 python -m pytest tests/test_ops_gpu.py -x -q -m gpu > gpurun_out/r2_t14_ops.log 2>&1; echo "ops rc=$?"; tail -3 gpurun_out/r2_t14_ops.log
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "durations or benched or two_sessions or latency_path" > gpurun_out/r2_t14_par.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/r2_t14_par.log
KKX_PROFILE_DETAIL=1 python tools/profile_step.py --batch 64 > gpurun_out/r2_step_b64_v14_detail.txt 2>&1
head -1 gpurun_out/r2_step_b64_v14_detail.txt; grep tf32x3 gpurun_out/r2_step_b64_v14_detail.txt | head -8
python tools/profile_step.py --batch 64 > gpurun_out/r2_step_b64_v14.txt 2>&1; head -8 gpurun_out/r2_step_b64_v14.txt
for d in 0 16 32 48; do
  KKX_LIB=kokorox_b200/lib/libkkx_exp.so KKX_ARB_DBG=$d KKX_PROFILE_DETAIL=1 timeout 300 python tools/profile_step.py --batch 64 2>&1 | grep -E "arb_conv\[c128 k(3|7|11) d1 conv[12] m5746720|arb_conv\[c256 k(3|7|11) d1 conv[12] m957780|launches_per_run" > gpurun_out/r2_arb_dbg${d}_v14.txt
  echo "dbg=$d"; cat gpurun_out/r2_arb_dbg${d}_v14.txt
done
KKX_LIB=kokorox_b200/lib/libkkx_timing.so KKX_ARB_TIMING=1 KKX_ARB_TIMING_KS=99 timeout 300 python tools/profile_step.py --batch 64 --runs 2 2>&1 | grep "tf32x3 timing" | tail -2 > gpurun_out/r2_gemm_roles_v14.txt
cat gpurun_out/r2_gemm_roles_v14.txt
