#!/bin/bash
# round 2: LSTM step evidence on the packed GEMM pipeline: launch list of one eager optimiser step, full captures of the 160-tile head GEMM
# (k_gemm_bf16x3<128, true>) and of the recurrence kernels, phase stamps and the A/B timings (each capture after its command exited 0 without ncu)
mkdir -p gpurun_out
L="python scripts/prof_lstm.py"
$L > gpurun_out/plain_lstm.log 2>&1 && {
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_lstm_step.csv $L > gpurun_out/ncu_l1.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_gemm_bf16x3 -s 18 -c 18 -f -o gpurun_out/r02_prof_gemm_lstm $L > gpurun_out/ncu_l2.log 2>&1
}
timeout 100 python scripts/r02/gemm_packing_ab.py > gpurun_out/r02_gemm_packing_ab.log 2>&1
timeout 100 python scripts/r02/gemm_stamps.py > gpurun_out/r02_gemm_stamps.log 2>&1
timeout 100 python scripts/lstm_stamps.py > gpurun_out/r02_lstm_stamps.log 2>&1
tail -n 2 gpurun_out/ncu_l1.log gpurun_out/ncu_l2.log; tail -n 4 gpurun_out/r02_gemm_packing_ab.log
