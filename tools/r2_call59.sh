timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "lstm" 2>&1 | tail -3
timeout 300 python tools/lstm_ab.py 2>&1 | tail -8
