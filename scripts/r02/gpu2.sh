#!/bin/bash
# round 2, call 2: clean-observation acting, MSE losses, resume tests, new bench contract (R chunks per step), reference arm, sweep
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 25 gpurun_out/pytest_gpu.log | cut -c1-400
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -n 4 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.log 2>&1; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"
timeout 900 python bench.py --sweep > gpurun_out/bench_sweep.log 2>&1; echo "sweep rc=$?"
tail -c 600 gpurun_out/bench_ref.log
