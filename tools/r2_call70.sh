timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "benched or golden or durations or batch" 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python tools/profile_step.py --batch 64 2>&1 | grep -E "gpu_ms|^adain_coef"
