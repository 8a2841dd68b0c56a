#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_lstm_gpu.py -m gpu -q -x > gpurun_out/pytest_lstm.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_lstm.log
tail -n 6 gpurun_out/pytest_lstm.log
timeout 200 python scripts/lstm_stamps.py 2>&1 | tail -20
