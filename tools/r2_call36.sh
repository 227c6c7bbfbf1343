set -x
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "attention" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "durations or benched or pair_gemm or two_sessions or latency_path" 2>&1 | tail -3
for i in 1 2; do timeout 300 python tools/profile_step.py --batch 64 2>&1 | grep -E "gpu_ms|^attention|^arb_conv |^conv_tc_tf32x3"; done
