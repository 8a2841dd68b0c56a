#!/bin/bash
# usage: bash scripts_gpu_tests.sh [pytest args...] ; then a quick bench
mkdir -p gpurun_out
timeout 900 python -m pytest ${@:-tests} -m gpu -q -x -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "err|parity|loss curve|passed|failed|rc=|Error|rollout" gpurun_out/pytest_gpu.log | tail -40
nvidia-smi --query-gpu=name --format=csv,noheader || echo "GPU UNRESPONSIVE"
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_quick.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_quick.log
tail -3 gpurun_out/bench_quick.log
