#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_lstm_gpu.py tests/test_dense_gpu.py -m gpu -q -x > gpurun_out/pytest_gemm.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gemm.log
tail -n 12 gpurun_out/pytest_gemm.log
timeout 200 python scripts/time_gemm.py 2>&1 | tail -13
timeout 200 python scripts/time_lstm.py 2>&1 | tail -4
