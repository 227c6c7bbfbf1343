set -x
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "bit_identical" > gpurun_out/r2_t16_pair.log 2>&1; rc=$?; echo "pair rc=$rc"; tail -15 gpurun_out/r2_t16_pair.log
if [ $rc -ne 0 ]; then nvidia-smi --query-gpu=name,memory.used --format=csv; exit 0; fi
timeout 900 python -m pytest tests/test_ops_gpu.py -x -q -m gpu > gpurun_out/r2_t16_ops.log 2>&1; echo "ops rc=$?"; tail -3 gpurun_out/r2_t16_ops.log
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "durations or benched or two_sessions or latency_path or composition or ragged" > gpurun_out/r2_t16_par.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/r2_t16_par.log
KKX_PROFILE_DETAIL=1 timeout 300 python tools/profile_step.py --batch 64 > gpurun_out/r2_step_b64_v16_detail.txt 2>&1
head -1 gpurun_out/r2_step_b64_v16_detail.txt; grep tf32x3 gpurun_out/r2_step_b64_v16_detail.txt | head -8
timeout 300 python tools/profile_step.py --batch 64 > gpurun_out/r2_step_b64_v16.txt 2>&1; head -6 gpurun_out/r2_step_b64_v16.txt
