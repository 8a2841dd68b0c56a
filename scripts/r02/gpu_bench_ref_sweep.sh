#!/bin/bash
# bench on the final code, reference arm, config-5 sweep on one GPU (+ the host-rollout tests)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_policy_gpu.py -m gpu -q -x > gpurun_out/pytest_policy.log 2>&1; echo "policy rc=$?"; tail -3 gpurun_out/pytest_policy.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_full.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"
timeout 900 python bench.py --sweep > gpurun_out/bench_sweep.log 2>&1; echo "sweep rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/bench_full.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value %.4g e2e %.4g sync %.4g u8 %.4g ret %.4g' % (d['value'], d['e2e']['value'], d['e2e']['synchronous_call']['value'], d['e2e']['done_as_u8']['value'], d['e2e_episode_returns']['value']))
        print('step_api', d['step_api']['value'], d['step_api']['roofline']['frac'])
PY
