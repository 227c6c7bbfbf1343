set -x
timeout 900 python -m pytest tests/test_ops_gpu.py -q -k "fused_arb" > gpurun_out/r2_t4_ops.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t4_ops.log
tail -25 gpurun_out/r2_t4_ops.log
timeout 900 python -m pytest tests/test_parity_gpu.py -q -s -k "bf16_residual_stream or concurrent_callers" > gpurun_out/r2_t4_par.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t4_par.log
grep -E "stream_bf16 stage|passed|failed|Error|assert" gpurun_out/r2_t4_par.log | head -20
cat > /tmp/opt_bench.py <<'PY'
import sys, json
sys.path.insert(0, '.')
from kokorox_b200.onn import B200Koko
from kokorox_b200.synth import ensure_weights, synth_batch
m = B200Koko.new(ensure_weights())
toks, styles, speeds = synth_batch(64, 510)
m.stage(toks, styles, speeds)
for name, opts in (("base", {}), ("stream_bf16", {"stream_bf16": 1})):
    for k in ("stream_bf16",):
        m.set_option(k, opts.get(k, 0))
    for _ in range(3):
        m.run_staged()
    m.profile_enable(True)
    m.run_staged()
    p = m.profile()
    m.profile_enable(False)
    kern = p["kernels"]
    top = {k: round(v[1] / 1e3, 2) for k, v in sorted(kern.items(), key=lambda kv: -kv[1][1])[:14]}
    print(json.dumps({"config": name, "gpu_ms": round(p["gpu_us"] / 1e3, 2), "arb_flops": p["arb_flops"], "arb_bytes": p["arb_bytes"], "kernels_ms": top}), flush=True)
m.close()
PY
python /tmp/opt_bench.py > gpurun_out/r2_opt_bench_v4.txt 2>&1
cat gpurun_out/r2_opt_bench_v4.txt
