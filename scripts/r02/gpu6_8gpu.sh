#!/bin/bash
# round 2, call 6 (8 GPUs): bench line at N = 8 (distill parity block, exchange phases, e2e at 8 ranks)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_8gpu.log 2> gpurun_out/bench_8gpu.err; echo "bench8 rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/bench_8gpu.log'):
    if l.startswith('{'):
        d=json.loads(l); print('value',d['value'],'e2e',d['e2e']['value']); print(json.dumps(d.get('distill'),indent=1)[:3500])
PY
