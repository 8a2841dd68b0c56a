#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_policy_gpu.py -m gpu -q -x > gpurun_out/pytest_policy.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_policy.log
tail -n 5 gpurun_out/pytest_policy.log
timeout 600 python scripts/e2e_sweep.py > gpurun_out/e2e_sweep.log 2>&1; echo "sweep rc=$?" >> gpurun_out/e2e_sweep.log
cat gpurun_out/e2e_sweep.log
