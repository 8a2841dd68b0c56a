#!/bin/bash
# round 2, call 12: rb_lstm2_step (graph), server with the four-output policy; full suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_lstm2_gpu.py tests/test_config1_gpu.py -m gpu -q -x > gpurun_out/pytest_l2.log 2>&1; echo "l2 rc=$?"; tail -5 gpurun_out/pytest_l2.log
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|rc=" gpurun_out/pytest_gpu.log | tail -3
python - <<'PY'
import torch, time
from reacherdistilation_b200.student_nn import StudentLSTM2, lstm2_spec
for B in (100, 2048):
    n2 = StudentLSTM2(spec=lstm2_spec(units=100, steps=20), seed=1)
    dev="cuda"
    o2, a2 = torch.randn((20, B, 11), device=dev), torch.randn((20, B, 2), device=dev) * 0.3
    t2 = torch.cat([torch.randn((20, B, 2), device=dev) * 0.3, -1 + 0.2 * torch.randn((20, B, 2), device=dev)], -1)
    r2 = torch.randn((20, B), device=dev) * 0.2
    for g in (False, True):
        f = (lambda: n2.step(o2, a2, t2, r2, None, keep_prob=0.5, seed=0)) if g else (lambda: (n2.loss_grad(o2, a2, t2, r2, None, keep_prob=0.5, seed=0, iteration=n2.t), n2.adam_step()))
        for _ in range(5): f()
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); e0.record()
        for _ in range(50): f()
        e1.record(); torch.cuda.synchronize()
        print("lstm2 B=%d graph=%s: %.3f ms per optimiser step" % (B, g, e0.elapsed_time(e1) / 50))
PY
