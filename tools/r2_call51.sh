set -x
KKX_LIB=kokorox_b200/lib/libkkx_exp2.so timeout 300 python tools/lstm_ab.py 2>&1 | grep -E "lstm G=|one set" | sort | uniq -c | sort -rn | head -40
