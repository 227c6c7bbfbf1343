set -x
timeout 900 python -m pytest tests/test_parity_gpu.py -q -x -k "phase_fused or bf16_batch_equals" > gpurun_out/r2_t8_par.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t8_par.log
tail -6 gpurun_out/r2_t8_par.log
python /dev/stdin > gpurun_out/r2_opt_bench_v8.txt 2>&1 <<'PY'
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
    ups = {k: [v[0], round(v[1] / 1e3, 3)] for k, v in kern.items() if " k2" in k}
    print(json.dumps({"config": name, "gpu_ms": round(p["gpu_us"] / 1e3, 2), "ups_total_ms": round(sum(v[1] for v in ups.values()), 3), "ups": ups}), flush=True)
m.close()
PY
cat gpurun_out/r2_opt_bench_v8.txt
