set -x
timeout 900 python -m pytest tests/test_ops_gpu.py -q -k "attention or split_fp16 or split_tf32" > gpurun_out/r2_t2_ops.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t2_ops.log
tail -30 gpurun_out/r2_t2_ops.log
timeout 900 python -m pytest tests/test_parity_gpu.py -q -k "round2_kernels or latency_path" > gpurun_out/r2_t2_par.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t2_par.log
tail -30 gpurun_out/r2_t2_par.log
cat > /tmp/opt_bench.py <<'PY'
import sys, json
sys.path.insert(0, '.')
from kokorox_b200.onn import B200Koko
from kokorox_b200.synth import ensure_weights, synth_batch
m = B200Koko.new(ensure_weights())
toks, styles, speeds = synth_batch(64, 510)
m.stage(toks, styles, speeds)
for name, opts in (("base", {}), ("attn_umma", {"attention_umma": 1}), ("split_f16", {"split_f16": 1}), ("both", {"attention_umma": 1, "split_f16": 1})):
    for k in ("attention_umma", "split_f16"):
        m.set_option(k, opts.get(k, 0))
    for _ in range(3):
        m.run_staged()
    m.profile_enable(True)
    m.run_staged()
    p = m.profile()
    m.profile_enable(False)
    kern = p["kernels"]
    top = {k: round(v[1] / 1e3, 2) for k, v in sorted(kern.items(), key=lambda kv: -kv[1][1])[:12]}
    print(json.dumps({"config": name, "gpu_ms": round(p["gpu_us"] / 1e3, 2), "kernels_ms": top}), flush=True)
m.close()
PY
python /tmp/opt_bench.py > gpurun_out/r2_opt_bench_v2.txt 2>&1
cat gpurun_out/r2_opt_bench_v2.txt
