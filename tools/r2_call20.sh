set -x
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "arb" 2>&1 | tail -2
KKX_PROFILE_DETAIL=1 timeout 300 python tools/profile_step.py --batch 64 > gpurun_out/r2_step_b64_v20_detail.txt 2>&1
head -1 gpurun_out/r2_step_b64_v20_detail.txt; grep "arb_conv\[" gpurun_out/r2_step_b64_v20_detail.txt | grep "m5746720\|m957780" 
timeout 300 python tools/profile_step.py --batch 64 2>&1 | head -4
