#!/bin/bash
# round 2, call 8: done-mask / return-sum outputs of the host rollout, limit force without division, lstm2 bench leg: tests, smoke, bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|rc=|Error|error" gpurun_out/pytest_gpu.log | tail -5
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -n 3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_full.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_full.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value %.4g e2e %.4g u8 %.4g ret %.4g full %.4g' % (d['value'], d['e2e']['value'], d['e2e']['done_as_u8']['value'], d['e2e_episode_returns']['value'], d['e2e_full_buffer']['value']))
        print('distill', d['distill']['value'], d['distill']['ms_per_step'], 'lstm', d['lstm']['ms_per_step'], 'lstm2', d['lstm2'])
        print('step_api', d['step_api']['value'], d['step_api']['roofline']['frac'], 'config1', d['config1'])
PY
