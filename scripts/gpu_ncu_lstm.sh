#!/bin/bash
# launch list of a few LSTM optimiser steps (no graph), after the same command exited 0 without ncu
mkdir -p gpurun_out
CMD="python scripts/prof_lstm.py"
$CMD > gpurun_out/plain_lstm.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_lstm.csv $CMD > gpurun_out/ncu_lstm.log 2>&1
tail -n 2 gpurun_out/plain_lstm.log gpurun_out/ncu_lstm.log
