set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"arb_conv_kernel" -s 30 -c 2 -f -o gpurun_out/r2_arb_k3_v10 python tools/profile_step.py --batch 64 --runs 1 > gpurun_out/ncu_arb_k3_v10.log 2>&1
tail -3 gpurun_out/ncu_arb_k3_v10.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm32p_kernel|attn_umma_kernel|attn_prep_kernel|apply_f16x2_kernel" -s 2 -c 6 -f -o gpurun_out/r2_gemm_attn_v10 python tools/profile_step.py --batch 64 --runs 1 > gpurun_out/ncu_gemm_attn_v10.log 2>&1
tail -3 gpurun_out/ncu_gemm_attn_v10.log
ls -la gpurun_out/*.ncu-rep | tail -3
