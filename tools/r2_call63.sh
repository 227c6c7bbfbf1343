set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2_bench_n2_v63.json 2> gpurun_out/r2_bench_n2_v63.err; echo "rc=$?"
tail -c 300 gpurun_out/r2_bench_n2_v63.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2_bench_n2_v63.json') if x.startswith('{')]
d=json.loads(l[-1])
print(d['value'], d['n_gpus'], d['ms_per_step'], d['e2e']['value'])
print(d.get('per_rank'))
print('cfg4', d['cfg4']['audio_s_per_s'] if d.get('cfg4') else None)
PY
