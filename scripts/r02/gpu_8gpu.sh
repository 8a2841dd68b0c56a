#!/bin/bash
# round 2, call 11 (8 GPUs): multi-GPU tests on the final code, bench line at N = 8, config-5 sweep at N = 8
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dp_fused_gpu.py -m gpu -q -x -s > gpurun_out/pytest_8gpu.log 2>&1; echo "pytest8 rc=$?"; grep -E "world|passed|failed" gpurun_out/pytest_8gpu.log | tail -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_8gpu.log 2> gpurun_out/bench_8gpu.err; echo "bench8 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --sweep > gpurun_out/bench_sweep_8gpu.log 2> gpurun_out/bench_sweep_8gpu.err; echo "sweep8 rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/bench_8gpu.log'):
    if l.startswith('{'):
        d=json.loads(l); print('value %.4g e2e %.4g ret %.4g' % (d['value'], d['e2e']['value'], d['e2e_episode_returns']['value'])); dd=d['distill']; print('distill %.4g %.2f us' % (dd['value'], dd['ms_per_step']*1e3), dd.get('parity',{}).get('ok'), dd.get('student_kernel_phases_us'))
PY
tail -c 700 gpurun_out/bench_sweep_8gpu.log
