#!/bin/bash
# round 2, call 5: resident env server (config 1), gpu-scope slab fence in the rollout kernel (e2e), full GPU suite, bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_config1_gpu.py tests/test_env_gpu.py -m gpu -q -x -s > gpurun_out/pytest_serve.log 2>&1; echo "serve rc=$?"; tail -5 gpurun_out/pytest_serve.log
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|rc=" gpurun_out/pytest_gpu.log | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/bench_full.log'):
    if l.startswith('{'):
        d=json.loads(l); print('value',d['value'],'e2e',d['e2e']['value'],'config1',d.get('config1'))
PY
