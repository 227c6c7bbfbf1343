set -x
for mf in 0 80000 0 80000; do
  timeout 300 python tools/profile_step.py --batch 64 --max-frames $mf 2>&1 | grep -E "gpu_ms|^arb_conv|^lstm|^conv_tc |Error|error" | head -6
done
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
