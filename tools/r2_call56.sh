set -x
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "pointwise or colstats or lstm" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "benched or bf16 or tensor_core or stream or batch" 2>&1 | tail -3
for v in 0 1 0 1; do KKX_PROFILE_DETAIL=1 timeout 300 python tools/profile_step.py --batch 64 --set fuse_noise_stats=$v 2>&1 | grep -E "gpu_ms|^colstats|^pointwise|ci22 co128|^apply_bf16"; done
