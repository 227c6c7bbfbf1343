timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "layernorm" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "durations or pair_gemm or benched or golden" 2>&1 | tail -3
for i in 1 2; do timeout 300 python tools/profile_step.py --batch 64 2>&1 | grep -E "gpu_ms|^layernorm"; done
timeout 300 python tools/profile_step.py --batch 1 2>&1 | grep -E "gpu_ms|^layernorm"
