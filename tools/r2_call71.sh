timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "golden or two_sessions or concurrent or lstm or benched" 2>&1 | tail -3
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "lstm or stft" 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
