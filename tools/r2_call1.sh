set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t1.log
tail -30 gpurun_out/r2_t1.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_v1.json 2> gpurun_out/r2_bench_v1.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2_bench_v1.err
python tools/profile_step.py --batch 1 --tokens 510 --runs 3 > gpurun_out/r2_step_b1_510_v1.txt 2>&1
python tools/profile_step.py --batch 1 --tokens 50 --runs 3 > gpurun_out/r2_step_b1_50_v1.txt 2>&1
KKX_PROFILE_DETAIL=1 python tools/profile_step.py --batch 64 > gpurun_out/r2_step_b64_v1_detail.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"istft_kernel" -c 2 -f -o gpurun_out/r2_istft_v1 python tools/profile_step.py --batch 8 --runs 1 > gpurun_out/ncu_istft_v1.log 2>&1
ls -la gpurun_out | tail -12
