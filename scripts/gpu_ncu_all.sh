#!/bin/bash
# full ncu captures of the three dominant kernels (rollout, student, LSTM GEMM), each after the same command exited 0 without ncu
mkdir -p gpurun_out
CMD="python bench.py --steps 5 --warmup 3 --quick"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_rollout_policy_tc -s 3 -c 1 -f -o gpurun_out/prof_rollout_tc $CMD > gpurun_out/ncu2.log 2>&1
CMD="python scripts/prof_distill.py"
$CMD > gpurun_out/plain_distill2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_student_tc -s 5 -c 1 -f -o gpurun_out/prof_student_tc $CMD > gpurun_out/ncu_d2.log 2>&1
CMD="python scripts/prof_lstm.py"
$CMD > gpurun_out/plain_lstm.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_gemm_bf16x3 -s 78 -c 1 -f -o gpurun_out/prof_lstm_gemm $CMD > gpurun_out/ncu_lstm2.log 2>&1
tail -n 2 gpurun_out/ncu2.log gpurun_out/ncu_d2.log gpurun_out/ncu_lstm2.log
