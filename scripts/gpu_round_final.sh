#!/bin/bash
# end-of-round evidence: full GPU test suite, smoke, full bench line, reference arm, config-5 sweep, launch lists + full ncu captures of the
# rollout and student kernels (each after its command exited 0 without ncu)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -n 3 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_full.log 2>&1; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"; tail -c 400 gpurun_out/bench_ref.log
timeout 600 python scripts/sweep_scaling.py > gpurun_out/sweep_1gpu.log 2>&1; echo "sweep rc=$?"
CMD="python bench.py --steps 5 --warmup 3 --quick"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_tc.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rollout_policy_tc -s 3 -c 1 -f -o gpurun_out/prof_rollout_tc $CMD > gpurun_out/ncu2.log 2>&1
CMD="python scripts/prof_distill.py"
$CMD > gpurun_out/plain_distill.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 30 --csv --log-file gpurun_out/launches_distill.csv $CMD > gpurun_out/ncu_d1.log 2>&1
$CMD > gpurun_out/plain_distill2.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_student_tc -s 5 -c 1 -f -o gpurun_out/prof_student_tc $CMD > gpurun_out/ncu_d2.log 2>&1
cat gpurun_out/plain_distill.log; tail -n 2 gpurun_out/ncu1.log gpurun_out/ncu2.log gpurun_out/ncu_d1.log gpurun_out/ncu_d2.log
nvidia-smi --query-gpu=name --format=csv,noheader || echo "GPU UNRESPONSIVE"
