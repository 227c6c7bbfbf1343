set -x
timeout 600 python -m pytest tests/test_ops_gpu.py -q -k "attention" > gpurun_out/r2_t7_ops.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t7_ops.log
tail -4 gpurun_out/r2_t7_ops.log
timeout 900 python -m pytest tests/test_parity_gpu.py -q -x -k "round2_kernels or durations_bit_exact or benched or two_sessions or latency_path" > gpurun_out/r2_t7_par.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t7_par.log
tail -4 gpurun_out/r2_t7_par.log
python tools/profile_step.py --batch 64 > gpurun_out/r2_step_b64_v7.txt 2>&1
head -14 gpurun_out/r2_step_b64_v7.txt
