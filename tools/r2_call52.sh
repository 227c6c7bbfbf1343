set -x
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "lstm" 2>&1 | tail -5
timeout 300 python tools/lstm_ab.py 2>&1 | tail -8
KKX_LIB=kokorox_b200/lib/libkkx_exp2.so timeout 300 python tools/lstm_ab.py 2>&1 | grep -E "lstm G=" | sort | uniq -c | sort -rn | awk '{ $1=""; print }' | sort -u -k3,5 | head -20
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "durations or benched or latency_path or smoke or reference_example" 2>&1 | tail -4
for v in 0 1; do timeout 300 python tools/profile_step.py --batch 64 --set lstm_fast_gates=$v 2>&1 | grep -E "gpu_ms|^lstm"; done
for v in 0 1; do timeout 300 python tools/profile_step.py --batch 1 --set lstm_fast_gates=$v 2>&1 | grep -E "gpu_ms|^lstm"; done
