set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t0.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t0.log
tail -3 gpurun_out/r2_t0.log
python tools/profile_step.py --batch 8 > gpurun_out/r2_step_b8_v0.txt 2>&1
for k in istft_kernel stft_kernel sine_source_kernel gather_rows_kernel colstats_kernel apply_bf16_kernel apply_tf32_kernel add_rows_kernel layernorm_kernel im2col_bf16_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^$k" -c 2 -f -o gpurun_out/r2_hbm_v0_$k python tools/profile_step.py --batch 8 --runs 1 > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out | tail -20
