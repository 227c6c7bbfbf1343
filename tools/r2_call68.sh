timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "im2col" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "benched or tensor_core or golden" 2>&1 | tail -3
for i in 1 2; do timeout 300 python tools/profile_step.py --batch 64 2>&1 | grep -E "gpu_ms|^im2col"; done
