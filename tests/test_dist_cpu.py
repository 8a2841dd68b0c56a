"""N>1 host logic on CPU with gloo, world_size 2: env sharding by global id and the gradient all-reduce reproduce the
single-process result on the concatenated batch (what the NCCL path does on the GPUs)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import nn_np as NN
from oracle import reacher_np as RN
from reacherdistilation_b200.dist import allreduce_gradloss, max_over_ranks, shard_range


def test_shard_range_partitions_exactly():
    for total in (1, 7, 8, 262144, 1000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(total, rank, world)
    env = RN.ReacherOracle(hi - lo, seed=3, env_offset=lo)
    obs = env.reset()
    rng = np.random.default_rng(0)                      # same student / teacher on every rank (MpiAdam.sync)
    P = (rng.standard_normal(NN.mlp_param_count()) * 0.1).astype(np.float32)
    x = np.concatenate([obs, np.zeros((hi - lo, 5))], -1)
    t = np.tile(np.array([0.1, -0.2, -1.0, -1.2]), (hi - lo, 1))
    s, hs = NN.mlp_fwd(x, P)
    l, ds = NN.kl_loss(s, t)
    gl = torch.from_numpy(np.concatenate([NN.mlp_bwd(hs, P, ds), [l]]))
    allreduce_gradloss(gl)
    tmax = max_over_ranks(float(rank + 1))
    from reacherdistilation_b200.dist import all_ranks_agree, rank_checkpoint_path
    assert all_ranks_agree(True) is True and all_ranks_agree(rank == 0) is False and all_ranks_agree(False) is False
    assert rank_checkpoint_path("a/b.pt", rank, world) == "a/b.pt.rank%d" % rank and rank_checkpoint_path("a/b.pt", 0, 1) == "a/b.pt"
    if rank == 0:
        q.put((gl.numpy(), obs, tmax))
    else:
        q.put((None, obs, tmax))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_allreduce_equals_concatenated_batch():
    total, world = 101, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in range(world)]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    gl = [r[0] for r in res if r[0] is not None][0]
    assert all(r[2] == 2.0 for r in res)
    # single-process reference on all 101 envs
    env = RN.ReacherOracle(total, seed=3)
    obs = env.reset()
    rng = np.random.default_rng(0)
    P = (rng.standard_normal(NN.mlp_param_count()) * 0.1).astype(np.float32)
    x = np.concatenate([obs, np.zeros((total, 5))], -1)
    t = np.tile(np.array([0.1, -0.2, -1.0, -1.2]), (total, 1))
    s, hs = NN.mlp_fwd(x, P)
    l, ds = NN.kl_loss(s, t)
    ref = np.concatenate([NN.mlp_bwd(hs, P, ds), [l]])
    assert np.allclose(gl, ref, rtol=1e-10, atol=1e-10)
    shard_obs = sorted([r[1] for r in res], key=lambda o: -o.shape[0])    # rank 0 has the larger shard (51)
    assert np.array_equal(np.concatenate(shard_obs), obs)
