set -x
CMD="python tools/profile_step.py --batch 16 --runs 1"
timeout 300 $CMD > gpurun_out/plain72.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv_phase_resident|pointwise_conv_stats" -c 4 -o gpurun_out/r2_new_kernels_a_v72 $CMD > gpurun_out/ncu72a.log 2>&1; echo "rc a=$?"
timeout 400 ncu --set full --clock-control none -k regex:"layernorm_vec|im2col_bf16_fast" -s 6 -c 3 -o gpurun_out/r2_new_kernels_b_v72 $CMD > gpurun_out/ncu72b.log 2>&1; echo "rc b=$?"
ls -la gpurun_out/*.ncu-rep
