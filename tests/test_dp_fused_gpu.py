"""Data-parallel student step with the gradient exchange fused into the kernel (one-shot all-reduce over NVLink peer memory) vs the
NCCL all-reduce path, on 2 GPUs (skipped on a single-GPU box): identical loss curves and parameters, all ranks bit-identical."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from reacherdistilation_b200 import MODE_TC, STUDENT_MLP
    from reacherdistilation_b200.mlp_train import DaggerTrainer
    out = {}
    for name, fused in (("nccl", False), ("fused", True)):
        tr = DaggerTrainer(num_envs=640, seed=3, device=rank, student_kind=STUDENT_MLP, mode=MODE_TC, env_offset=rank * 640, lr=1e-3,
                           fused_allreduce=fused)
        tr.sync_params()
        losses = []
        for _ in range(25):
            tr.step()
            posted = tr.wait_loss() if tr.use_graph else None          # graph path: the loss mailbox (host polls mapped memory)
            losses.append(float(tr.last_loss()))
            assert posted is None or np.float32(posted) == np.float32(losses[-1]), (posted, losses[-1])
        out[name] = (np.array(losses), tr.student.params.cpu().numpy().copy())
        tr.close()
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_fused_peer_allreduce_matches_nccl():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=300) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for name in ("nccl", "fused"):
        assert np.array_equal(res[0][name][0], res[1][name][0]) and np.array_equal(res[0][name][1], res[1][name][1]), name   # ranks agree
    ln, pn = res[0]["nccl"]
    lf, pf = res[0]["fused"]
    assert lf[-1] < lf[0]
    assert np.allclose(lf, ln, rtol=1e-5) and np.abs(pf - pn).max() <= 1e-5      # 2 ranks: a + b is the same sum in both paths
