#!/bin/bash
# round evidence: full GPU test suite, smoke, full bench line, config-5 sweep, launch lists + full ncu captures (each after its command exited 0 without ncu)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -n 3 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_full.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_full.log
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?" >> gpurun_out/bench_ref.log; tail -c 600 gpurun_out/bench_ref.log
timeout 600 python scripts/sweep_scaling.py > gpurun_out/sweep_1gpu.log 2>&1; echo "sweep rc=$?" >> gpurun_out/sweep_1gpu.log
bash scripts/gpu_ncu.sh > /dev/null 2>&1
bash scripts/gpu_ncu_all.sh > /dev/null 2>&1
CMD="python scripts/prof_distill.py"
$CMD > gpurun_out/plain_distill.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/launches_distill.csv $CMD > gpurun_out/ncu_d1.log 2>&1
bash scripts/gpu_ncu_lstm.sh > /dev/null 2>&1
tail -n 2 gpurun_out/ncu2.log gpurun_out/ncu_d2.log gpurun_out/ncu_lstm2.log
tail -c 300 gpurun_out/bench_full.log
