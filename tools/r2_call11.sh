set -x
timeout 600 python -m pytest tests/test_ops_gpu.py -q -k "fused_arb" > gpurun_out/r2_t11_ops.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t11_ops.log
tail -4 gpurun_out/r2_t11_ops.log
timeout 900 python -m pytest tests/test_parity_gpu.py -q -x -k "bf16_tensor_core or bf16_batch_equals or many_short or benched" > gpurun_out/r2_t11_par.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t11_par.log
tail -4 gpurun_out/r2_t11_par.log
KKX_PROFILE_DETAIL=1 python tools/profile_step.py --batch 64 > gpurun_out/r2_step_b64_v11_detail.txt 2>&1
head -3 gpurun_out/r2_step_b64_v11_detail.txt; grep "arb_conv" gpurun_out/r2_step_b64_v11_detail.txt | head -24
