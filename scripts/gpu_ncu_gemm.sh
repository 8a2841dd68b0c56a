#!/bin/bash
mkdir -p gpurun_out
CMD="python scripts/prof_lstm.py"
$CMD > gpurun_out/plain_lstm.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_gemm_bf16x3 -s ${1:-78} -c 1 -f -o gpurun_out/prof_lstm_gemm $CMD > gpurun_out/ncu_lstm2.log 2>&1
tail -n 2 gpurun_out/ncu_lstm2.log
