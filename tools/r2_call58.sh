set -x
timeout 300 python tools/lstm_ab.py > gpurun_out/lstm_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_cluster -s 3 -c 1 -o gpurun_out/r2_lstm_v58 python tools/lstm_ab.py > gpurun_out/lstm_ncu.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/lstm_ncu.log
