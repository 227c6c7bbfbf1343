set -x
timeout 600 python bench.py --config 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain64.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 2200 --csv --log-file gpurun_out/r2_launches_bench_v64.csv python bench.py --config 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_v64.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/r2_launches_bench_v64.csv
