"""e2e (host-buffer C-ABI call) variants of the config-3 rollout: zero-copy result vs slab copies, slab count / first-slab length.
One subprocess per setting (the knobs are read once per process).  Usage: python scripts/e2e_sweep.py  ->  one JSON line per setting."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child():
    sys.path.insert(0, ROOT)
    import torch
    from reacherdistilation_b200 import MODE_TC
    from reacherdistilation_b200.env import VecReacher
    from reacherdistilation_b200.teacher import init_policy_params
    n, T, K = 65536, 50, 20
    env = VecReacher(num_envs=n, seed=0)
    env.reset()
    p = torch.from_numpy(init_policy_params(seed=0)).pin_memory()
    out = dict(obs=None, pdflat=None, rew=torch.empty((T, n)).pin_memory(), done=torch.empty((T, n), dtype=torch.uint8).pin_memory())
    if os.environ.get("E2E_FULL") == "1":
        out["obs"], out["pdflat"] = torch.empty((T, n, 11)).pin_memory(), torch.empty((T, n, 4)).pin_memory()
    fn = lambda: env.rollout_policy_host(p, T, nout=2, mode=MODE_TC, out=out)
    for _ in range(3):
        fn()
    best = []
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(K):
            fn()
        best.append((time.perf_counter() - t0) / K)
    print(json.dumps(dict(knobs={k: v for k, v in os.environ.items() if k.startswith("RB_HOST") or k == "E2E_FULL"}, ms_per_call=[round(1e3 * b, 4) for b in best],
                          env_steps_per_s=n * T / min(best), mean_rew=float(out["rew"].mean()))), flush=True)


if __name__ == "__main__":
    if os.environ.get("E2E_CHILD") == "1":
        child()
        sys.exit(0)
    settings = [dict(),                                                                       # defaults: reward kernel-written, done copied per in-kernel progress slab
                dict(RB_HOST_PROGRESS_SLABS="3"), dict(RB_HOST_PROGRESS_SLABS="10"),
                dict(RB_HOST_ZEROCOPY="0"), dict(RB_HOST_ZEROCOPY="0", RB_HOST_PROGRESS_SLABS="10"),   # everything by the copy engine, progress slabs
                dict(RB_HOST_ZEROCOPY="3"),                                                   # reward and done kernel-written
                dict(RB_HOST_PROGRESS="0"),                                                   # previous default: reward kernel-written, done copied after the kernel
                dict(RB_HOST_PROGRESS="0", RB_HOST_ZEROCOPY="0", RB_HOST_SLABS="5", RB_HOST_SLAB_FIRST="100"),      # the first schedule: 5 kernel slabs, all copied
                dict(E2E_FULL="1"), dict(E2E_FULL="1", RB_HOST_PROGRESS_SLABS="10"),
                dict(RB_HOST_PROGRESS="0", E2E_FULL="1")]
    for s in settings:
        envv = dict(os.environ, E2E_CHILD="1", **s)
        subprocess.run([sys.executable, os.path.abspath(__file__)], env=envv, check=False)
