#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dagger_gpu.py tests/test_student_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x > gpurun_out/pytest_dagger.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_dagger.log
tail -n 5 gpurun_out/pytest_dagger.log
cat > /tmp/ab.py <<'EOF'
import os, sys, time, torch
sys.path.insert(0, os.getcwd())
from reacherdistilation_b200 import MODE_TC
from reacherdistilation_b200.mlp_train import DaggerTrainer
for n in (32768, 262144):
    tr = DaggerTrainer(num_envs=n, seed=0, mode=MODE_TC)
    for _ in range(10): tr.step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(200): tr.step()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 200
    t0 = time.perf_counter()
    for _ in range(200): tr.step(); tr.wait_loss()
    e2e = (time.perf_counter() - t0) / 200 * 1e6
    import ctypes
    from reacherdistilation_b200._lib import lib
    buf = (ctypes.c_ulonglong * 48)()
    lib().rb_debug_student_timers(buf)
    ts = [buf[i] for i in range(12)]
    names = ["fold+image", "sync", "image load", "tiles", "dump+act", "sync", "reduce", "sync", "un-fold", "sync+adam", "teardown"]
    print("   CTA0 phases (us): " + ", ".join("%s %.1f" % (nm, (ts[i + 1] - ts[i]) / 1e3) for i, nm in enumerate(names)) + "  total %.1f" % ((ts[11] - ts[0]) / 1e3))
    print("fuse=%s envs=%d  %.2f us/iter  %.3e samples/s   e2e %.2f us  loss %.4f" % (os.environ.get("RB_DAGGER_FUSE_ACT", "1"), n, us, n / us * 1e6, e2e, tr.wait_loss()), flush=True)
    tr.close()
EOF
RB_DAGGER_FUSE_ACT=1 python /tmp/ab.py
RB_DAGGER_FUSE_ACT=0 python /tmp/ab.py
