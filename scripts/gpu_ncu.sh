#!/bin/bash
# launch list + one full capture of the tensor-core rollout kernel (bench --quick = headline section only)
mkdir -p gpurun_out
CMD="python bench.py --steps 5 --warmup 3 --quick"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_tc.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_rollout_policy_tc -s 3 -c 1 -f -o gpurun_out/prof_rollout_tc $CMD > gpurun_out/ncu2.log 2>&1
tail -n 3 gpurun_out/ncu1.log gpurun_out/ncu2.log
