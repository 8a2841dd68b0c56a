#!/bin/bash
mkdir -p gpurun_out
CMD="python scripts/prof_lstm_sizes.py"
timeout 200 $CMD > gpurun_out/plain_lstm_sizes.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_lstm_recur --csv --log-file gpurun_out/launches_lstm_sizes.csv $CMD > gpurun_out/ncu_lstm_sizes.log 2>&1
cat gpurun_out/plain_lstm_sizes.log
grep recur gpurun_out/launches_lstm_sizes.csv | awk -F'","' '{print $5, $9, $15}' | head -40
