timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "phase_fused" 2>&1 | tail -3
for v in 1 3 1 3; do KKX_PROFILE_DETAIL=1 timeout 300 python tools/profile_step.py --batch 64 --set ups_phase_loop=$v 2>&1 | grep -E "gpu_ms|co128 k2 phases"; done
