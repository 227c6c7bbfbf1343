set -x
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2_v9.json 2> gpurun_out/r2_bench_n2_v9.err; echo "bench n2 rc=$?"
tail -c 800 gpurun_out/r2_bench_n2_v9.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2_bench_n2_v9.json') if x.startswith('{')]
d=json.loads(l[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['n_gpus'])
print(json.dumps(d['per_rank']))
print(json.dumps(d['cfg4'])[:1200])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_bench_ref_n2_v9.json 2>&1; echo "ref rc=$?"; tail -c 600 gpurun_out/r2_bench_ref_n2_v9.json
