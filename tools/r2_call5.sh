set -x
timeout 900 python -m pytest tests/test_ops_gpu.py -q > gpurun_out/r2_t5_ops.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t5_ops.log
tail -8 gpurun_out/r2_t5_ops.log
timeout 1200 python -m pytest tests/test_parity_gpu.py tests/test_onnx_loader.py -m gpu -q -x > gpurun_out/r2_t5_par.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t5_par.log
tail -8 gpurun_out/r2_t5_par.log
KKX_PROFILE_DETAIL=1 python tools/profile_step.py --batch 64 > gpurun_out/r2_step_b64_v5_detail.txt 2>&1
head -30 gpurun_out/r2_step_b64_v5_detail.txt
python tools/profile_step.py --batch 1 --tokens 510 --runs 3 > gpurun_out/r2_step_b1_510_v5.txt 2>&1
python tools/profile_step.py --batch 1 --tokens 50 --runs 3 > gpurun_out/r2_step_b1_50_v5.txt 2>&1
head -3 gpurun_out/r2_step_b1_510_v5.txt gpurun_out/r2_step_b1_50_v5.txt
cat > /tmp/sb_detail.py <<'PY'
import sys, json, os
os.environ["KKX_PROFILE_DETAIL"] = "1"
sys.path.insert(0, '.')
from kokorox_b200.onn import B200Koko
from kokorox_b200.synth import ensure_weights, synth_batch
m = B200Koko.new(ensure_weights())
toks, styles, speeds = synth_batch(64, 510)
m.stage(toks, styles, speeds)
m.set_option("stream_bf16", 1)
for _ in range(3):
    m.run_staged()
m.profile_enable(True)
m.run_staged()
p = m.profile()
kern = p["kernels"]
print(json.dumps({"gpu_ms": p["gpu_us"] / 1e3}))
for k, v in sorted(kern.items(), key=lambda kv: -kv[1][1]):
    if k.startswith("arb_conv"):
        print(f"{k:48s} n={v[0]:4d} total {v[1]/1e3:9.3f} ms")
m.close()
PY
python /tmp/sb_detail.py > gpurun_out/r2_stream_bf16_detail_v5.txt 2>&1
head -40 gpurun_out/r2_stream_bf16_detail_v5.txt
