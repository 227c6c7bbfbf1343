set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8_v65.json 2> gpurun_out/r2_bench_n8_v65.err; echo "rc=$?"
tail -c 300 gpurun_out/r2_bench_n8_v65.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2_bench_n8_v65.json') if x.startswith('{')]
d=json.loads(l[-1])
print(d['value'], d['n_gpus'], d['ms_per_step'], d['e2e']['value'])
for r in d.get('per_rank', []): print(r)
print('cfg4', d['cfg4']['audio_s_per_s'] if d.get('cfg4') else None)
PY
