#!/bin/bash
# round 2, call 4: lr_t of the step computed by k_student_image (off the cooperative kernel setup path)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|rc=" gpurun_out/pytest_gpu.log | tail -3
timeout 300 python scripts/r02/student_stamps.py 32768 0.5 > gpurun_out/stamps_32k.log 2>&1; cat gpurun_out/stamps_32k.log
timeout 300 python scripts/r02/student_stamps.py 32768 1.0 > gpurun_out/stamps_32k_kp1.log 2>&1; cat gpurun_out/stamps_32k_kp1.log
timeout 300 python scripts/r02/student_stamps.py 262144 0.5 > gpurun_out/stamps_256k.log 2>&1; cat gpurun_out/stamps_256k.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.log 2>&1; echo "bench rc=$?"
