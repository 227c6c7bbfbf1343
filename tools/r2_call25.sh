set -x
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "pair_gemm or durations or benched or latency_path or two_sessions" 2>&1 | tail -5
for p in 1 0 1 0; do
  timeout 300 python tools/profile_step.py --batch 64 --set fuse_planes=$p 2>&1 | grep -E "gpu_ms|conv_tc_tf32x3|apply_f16x2|layernorm|attn_prep|attention" 
done
