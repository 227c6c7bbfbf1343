set -x
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "lstm" 2>&1 | tail -5
timeout 300 python tools/lstm_ab.py 2>&1 | tail -8
for pp in 0 1; do timeout 300 python tools/profile_step.py --batch 64 --set lstm_pingpong=$pp 2>&1 | grep -E "gpu_ms|^lstm"; done
