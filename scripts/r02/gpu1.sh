#!/bin/bash
# round 2, call 1: new physics on the GPU -- full GPU suite, smoke, bench line, instruction counts of k_step and the fused rollout
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -n 4 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.log 2>&1; echo "bench rc=$?"
python scripts/prof_step.py > gpurun_out/plain_step.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_xu.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_step -s 10 -c 2 --csv --log-file gpurun_out/ncu_step_counts.csv python scripts/prof_step.py > gpurun_out/ncu_step.log 2>&1
CMD="python bench.py --steps 5 --warmup 3 --quick"
$CMD > gpurun_out/plain_quick.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_xu.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_rollout_policy_tc -s 3 -c 2 --csv --log-file gpurun_out/ncu_rollout_counts.csv $CMD > gpurun_out/ncu_rollout.log 2>&1
tail -n 3 gpurun_out/plain_step.log gpurun_out/ncu_step.log gpurun_out/ncu_rollout.log
