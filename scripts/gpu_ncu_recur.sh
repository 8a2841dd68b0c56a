#!/bin/bash
# full ncu captures of the two LSTM recurrence kernels (after the same command exited 0 without ncu)
mkdir -p gpurun_out
CMD="python scripts/prof_lstm.py"
timeout 200 $CMD > gpurun_out/plain_recur.log 2>&1 &&
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_lstm_recur_fwd -s 2 -c 1 -f -o gpurun_out/prof_recur_fwd $CMD > gpurun_out/ncu_recur_f.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_lstm_recur_bwd -s 2 -c 1 -f -o gpurun_out/prof_recur_bwd $CMD > gpurun_out/ncu_recur_b.log 2>&1
tail -n 3 gpurun_out/ncu_recur_f.log gpurun_out/ncu_recur_b.log
