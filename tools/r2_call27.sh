set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm32p2_kernel|conv_pair_kernel" -s 4 -c 8 -f -o gpurun_out/r2_pair_kernels_v27 python tools/profile_step.py --batch 64 --runs 1 > gpurun_out/ncu_pair_v27.log 2>&1
tail -3 gpurun_out/ncu_pair_v27.log
ls -la gpurun_out/*.ncu-rep | tail -2
