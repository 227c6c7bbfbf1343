set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_pair_kernel" -s 2 -c 4 -f -o gpurun_out/r2_conv_pair_v35 python tools/profile_step.py --batch 64 --runs 1 > gpurun_out/ncu_conv_pair_v35.log 2>&1
tail -2 gpurun_out/ncu_conv_pair_v35.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"arb_conv_kernel" -s 30 -c 2 -f -o gpurun_out/r2_arb_k3_v35 python tools/profile_step.py --batch 64 --runs 1 > gpurun_out/ncu_arb_k3_v35.log 2>&1
tail -2 gpurun_out/ncu_arb_k3_v35.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm32p2_kernel" -s 4 -c 4 -f -o gpurun_out/r2_gemm_pair_v35 python tools/profile_step.py --batch 64 --runs 1 > gpurun_out/ncu_gemm_pair_v35.log 2>&1
tail -2 gpurun_out/ncu_gemm_pair_v35.log
ls -la gpurun_out/*v35.ncu-rep
