set -x
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "attention" 2>&1 | tail -6
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "durations or benched or pair_gemm or two_sessions or latency_path" 2>&1 | tail -4
for i in 1 2; do timeout 300 python tools/profile_step.py --batch 64 2>&1 | grep -E "gpu_ms|^attention|^attn_prep|^conv_tc_tf32x3"; done
KKX_LIB=kokorox_b200/lib/libkkx_timing.so KKX_ARB_TIMING=1 KKX_ARB_TIMING_KS=99 timeout 300 python tools/profile_step.py --batch 64 --runs 2 2>&1 | grep -E "attention timing" | tail -1
