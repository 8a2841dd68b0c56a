#!/bin/bash
# round 2, call 7: full GPU suite on the final code, then the ncu evidence of the round (each capture after its command exited 0 without ncu):
# launch lists of the bench step and the DAgger iteration, full captures of k_rollout_policy_tc / k_student_tc / k_step, counter CSVs bench.py reads
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|rc=" gpurun_out/pytest_gpu.log | tail -3
M1="gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_xu.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum"
Q="python bench.py --steps 2 --warmup 3 --quick"
$Q > gpurun_out/plain_quick.log 2>&1 && {
  ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_launches_rollout.csv $Q > gpurun_out/ncu_q1.log 2>&1
  ncu --metrics $M1 --clock-control none -k regex:k_rollout_policy_tc -s 3 -c 2 --csv --log-file gpurun_out/r02_ncu_counts_k_rollout_policy_tc.csv $Q > gpurun_out/ncu_q2.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_rollout_policy_tc -s 3 -c 1 -f -o gpurun_out/r02_prof_rollout_tc $Q > gpurun_out/ncu_q3.log 2>&1
}
D="python scripts/prof_distill.py"
$D > gpurun_out/plain_distill.log 2>&1 && {
  ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/r02_launches_distill_step.csv $D > gpurun_out/ncu_d1.log 2>&1
  ncu --metrics $M1 --clock-control none -k regex:k_student_tc -s 5 -c 2 --csv --log-file gpurun_out/r02_ncu_counts_k_student_tc.csv $D > gpurun_out/ncu_d2.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_student_tc -s 5 -c 1 -f -o gpurun_out/r02_prof_student_tc $D > gpurun_out/ncu_d3.log 2>&1
}
S="python scripts/prof_step.py"
$S > gpurun_out/plain_step.log 2>&1 && {
  ncu --metrics $M1 --clock-control none -k regex:k_step -s 10 -c 2 --csv --log-file gpurun_out/r02_ncu_counts_k_step.csv $S > gpurun_out/ncu_s1.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_step -s 10 -c 1 -f -o gpurun_out/r02_prof_k_step $S > gpurun_out/ncu_s2.log 2>&1
}
cat gpurun_out/plain_distill.log gpurun_out/plain_step.log; tail -n 2 gpurun_out/ncu_q3.log gpurun_out/ncu_d3.log gpurun_out/ncu_s2.log
ls -la gpurun_out/*.ncu-rep
