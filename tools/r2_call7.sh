set -x
timeout 600 python -m pytest tests/test_ops_gpu.py -q -k "attention" > gpurun_out/r2_t7_ops.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t7_ops.log
tail -4 gpurun_out/r2_t7_ops.log
timeout 900 python -m pytest tests/test_parity_gpu.py -q -x -k "round2_kernels or durations_bit_exact or benched or two_sessions or latency_path or phase_fused" > gpurun_out/r2_t7_par.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t7_par.log
tail -4 gpurun_out/r2_t7_par.log
cat > /tmp/opt_bench.py <<'PY'
import sys, json, os
os.environ["KKX_PROFILE_DETAIL"] = "1"
sys.path.insert(0, '.')
from kokorox_b200.onn import B200Koko
from kokorox_b200.synth import ensure_weights, synth_batch
m = B200Koko.new(ensure_weights())
toks, styles, speeds = synth_batch(64, 510)
m.stage(toks, styles, speeds)
for name, opts in (("base", {}), ("fuse_phases", {"fuse_phases": 1})):
    m.set_option("fuse_phases", opts.get("fuse_phases", 0))
    for _ in range(3):
        m.run_staged()
    m.profile_enable(True)
    m.run_staged()
    p = m.profile()
    m.profile_enable(False)
    kern = p["kernels"]
    ups = {k: round(v[1] / 1e3, 3) for k, v in kern.items() if "k2" in k}
    att = {k: round(v[1] / 1e3, 3) for k, v in kern.items() if k.startswith("att")}
    print(json.dumps({"config": name, "gpu_ms": round(p["gpu_us"] / 1e3, 2), "ups": ups, "attention": att}), flush=True)
m.close()
PY
python /tmp/opt_bench.py > gpurun_out/r2_opt_bench_v7.txt 2>&1
cat gpurun_out/r2_opt_bench_v7.txt
