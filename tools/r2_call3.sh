set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t3.log
tail -15 gpurun_out/r2_t3.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_v3.json 2> gpurun_out/r2_bench_v3.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_bench_v3.err
KKX_PROFILE_DETAIL=1 python tools/profile_step.py --batch 64 > gpurun_out/r2_step_b64_v3_detail.txt 2>&1
python tools/profile_step.py --batch 1 --tokens 510 --runs 3 > gpurun_out/r2_step_b1_510_v3.txt 2>&1
python tools/profile_step.py --batch 1 --tokens 50 --runs 3 > gpurun_out/r2_step_b1_50_v3.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_v3.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke_v3.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"attn_umma_kernel|gemm32p_kernel" -c 4 -f -o gpurun_out/r2_attn_gemm_v3 python tools/profile_step.py --batch 64 --runs 1 > gpurun_out/ncu_attn_gemm_v3.log 2>&1
ls -la gpurun_out | tail -8
