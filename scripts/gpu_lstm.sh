#!/bin/bash
# GEMM + LSTM parity tests, LSTM step timing (graph) and the launch list of one un-graphed step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_lstm_gpu.py -m gpu -q -x > gpurun_out/pytest_lstm.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_lstm.log
tail -n 4 gpurun_out/pytest_lstm.log
timeout 300 python scripts/time_lstm.py 2>&1 | tail -5
bash scripts/gpu_ncu_lstm.sh > /dev/null 2>&1
