set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t69.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t69.log
tail -5 gpurun_out/r2_t69.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_v69.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_smoke_v69.log
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_v69.json 2> gpurun_out/r2_bench_v69.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r2_bench_v69.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2_bench_v69.json') if x.startswith('{')]
d=json.loads(l[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
print(json.dumps(d['latency'])[:500])
print(d['roofline']['frac'], d['roofline']['hbm_frac'], d['roofline']['kernel_ms'])
print('cfg3', d['cfg3']['headline'] if d.get('cfg3') else None)
print('cfg4', d['cfg4']['audio_s_per_s'] if d.get('cfg4') else None)
print('cpu', d.get('cpu_baseline'))
PY
KKX_PROFILE_DETAIL=1 python tools/profile_step.py --batch 64 > gpurun_out/r2_step_b64_v69_detail.txt 2>&1
python tools/profile_step.py --batch 1 > gpurun_out/r2_step_b1_510_v69.txt 2>&1
python tools/profile_step.py --batch 1 --tokens 50 > gpurun_out/r2_step_b1_50_v69.txt 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_v69.json 2>&1; echo "ref rc=$?"
python tools/profile_step.py --batch 64 > gpurun_out/r2_step_b64_v69.txt 2>&1
