#!/bin/bash
# LSTM parity tests + optimiser-step timing with the persistent recurrence kernels (default) and with the GEMM + cell launch sequence
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_lstm_gpu.py -m gpu -q -x > gpurun_out/pytest_lstm.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_lstm.log
tail -n 15 gpurun_out/pytest_lstm.log
echo "--- recurrence kernels"; timeout 120 python scripts/time_lstm.py 2>&1 | tail -5
echo "--- launch sequence"; RB_LSTM_RECUR=0 timeout 120 python scripts/time_lstm.py 2>&1 | tail -5
