set -x
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "lstm" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "benched or composition or ragged or pair_gemm" 2>&1 | tail -2
for i in 1 2; do timeout 300 python tools/profile_step.py --batch 64 2>&1 | grep -E "gpu_ms|^lstm|^attention"; done
