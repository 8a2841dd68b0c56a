#!/bin/bash
# end-of-round refresh after the LSTM graph change: full GPU suite, smoke, bench line, LSTM launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -n 2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_full.log 2>&1; echo "bench rc=$?"
CMD="python scripts/prof_lstm.py"
$CMD > gpurun_out/plain_lstm.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_lstm.csv $CMD > gpurun_out/ncu_lstm.log 2>&1
tail -n 2 gpurun_out/plain_lstm.log gpurun_out/ncu_lstm.log
