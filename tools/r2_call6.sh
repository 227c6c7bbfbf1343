set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t6.log
tail -6 gpurun_out/r2_t6.log
KKX_LIB=$PWD/kokorox_b200/lib/libkkx_timing.so KKX_ARB_TIMING=1 python tools/profile_step.py --batch 64 --runs 2 > gpurun_out/r2_gemm_timing_v6.txt 2>&1
grep "tf32x3 timing" gpurun_out/r2_gemm_timing_v6.txt | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 --config 2 > gpurun_out/r2_bench_v6.json 2> gpurun_out/r2_bench_v6.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2_bench_v6.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2_bench_v6.json') if x.startswith('{')]
d=json.loads(l[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['latency'])
PY
