for rep in 1 2; do
for l in libkkx libkkx_s0t1 libkkx_s1t0 libkkx_s1t1; do echo "== $l"; KKX_LIB=kokorox_b200/lib/$l.so timeout 300 python tools/lstm_ab.py 2>&1 | tail -5 | cut -c1-100; done
done
