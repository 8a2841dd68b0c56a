"""e2e transport at N ranks (torchrun): rb_env_rollout_policy_host per-call time, max over ranks, for
   default | pinned buffers first-touched on the GPU's NUMA node | reward through the copy engine | both."""
import os, sys, time, glob
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from reacherdistilation_b200 import MODE_TC, _lib
from reacherdistilation_b200.dist import init_from_env, max_over_ranks
from reacherdistilation_b200.env import VecReacher
from reacherdistilation_b200.teacher import init_policy_params
rank, world, local = init_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
n, T = 65536, 50
def gpu_numa(i):
    p = torch.cuda.get_device_properties(i)
    bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
    try:
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
    except Exception as e:
        node = -1
    return bdf, node
def node_cpus(node):
    try:
        txt = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
    except Exception:
        return None
    cpus = []
    for part in txt.split(","):
        a, _, b = part.partition("-")
        cpus += list(range(int(a), int(b or a) + 1))
    return cpus
bdf, node = gpu_numa(local)
aff0 = sorted(os.sched_getaffinity(0))
print("rank %d gpu %s numa %d, affinity %d cpus [%d..%d], nodes %s" % (rank, bdf, node, len(aff0), aff0[0], aff0[-1], sorted(glob.glob("/sys/devices/system/node/node*"))), flush=True)
p_host = torch.from_numpy(init_policy_params(seed=0)).pin_memory()
def run(label, bind, transport, mask=False):
    if bind and node >= 0:
        cpus = [c for c in (node_cpus(node) or []) if c in aff0]
        if cpus:
            os.sched_setaffinity(0, cpus)
    rew = torch.empty((T, n)).pin_memory(); done = torch.empty((T, n), dtype=torch.uint8).pin_memory()
    rew.zero_(); done.zero_()                     # first touch
    env = VecReacher(num_envs=n, seed=0, device=local, env_offset=rank * n)
    env.reset()
    _lib.check(_lib.lib().rb_env_set_host_transport(env._h, transport))
    out = dict(obs=None, pdflat=None, rew=rew, done=done)
    if mask:
        out = dict(obs=None, pdflat=None, rew=rew, done=None, done_mask=torch.zeros((n,), dtype=torch.int64).pin_memory())
    fn = lambda: env.rollout_policy_host(p_host, T, nout=2, mode=MODE_TC, out=out)
    for _ in range(5): fn()
    if world > 1: torch.distributed.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(100): fn()
    dt = time.perf_counter() - t0
    if world > 1: torch.distributed.barrier()
    ms = max_over_ranks(dt / 100 * 1e3, dev)
    if rank == 0:
        print("%-40s %.3f ms per call (max over %d ranks) = %.3g env-steps/s total" % (label, ms, world, n * T * world / (ms * 1e-3)), flush=True)
    env.close()
    os.sched_setaffinity(0, aff0)
run("default (kernel-stored reward)", False, 1)
run("copy engine for reward and done", False, 0)
run("kernel-stored reward + done", False, 3)
run("kernel-stored reward + done MASK (8 B per env)", False, 1, True)
run("copy-engine reward + kernel-stored done mask", False, 0, True)
