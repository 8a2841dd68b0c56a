#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_student_gpu.py tests/test_dagger_gpu.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 5 gpurun_out/pytest_gpu.log
CMD="python scripts/prof_distill.py $1"
$CMD > gpurun_out/plain_distill.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/launches_distill.csv $CMD > gpurun_out/ncu_d1.log 2>&1
$CMD > gpurun_out/plain_distill2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_student_tc -s 5 -c 1 -f -o gpurun_out/prof_student_tc $CMD > gpurun_out/ncu_d2.log 2>&1
cat gpurun_out/plain_distill.log; tail -n 3 gpurun_out/ncu_d1.log gpurun_out/ncu_d2.log
