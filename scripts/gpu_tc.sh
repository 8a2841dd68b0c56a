#!/bin/bash
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_policy_gpu.py -m gpu -q -s -x > gpurun_out/pytest_tc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tc.log
grep -E "err|rollout|passed|failed|rc=|Error|error" gpurun_out/pytest_tc.log | tail -30
nvidia-smi --query-gpu=name --format=csv,noheader || echo "GPU UNRESPONSIVE"
timeout 300 python bench.py --steps 50 --warmup 5 --quick > gpurun_out/bench_tc.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_tc.log
tail -3 gpurun_out/bench_tc.log
