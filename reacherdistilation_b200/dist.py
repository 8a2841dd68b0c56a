"""Multi-GPU plumbing: one process per GPU (torchrun), env shards by contiguous GLOBAL env id, one sum all-reduce of the
flat [grad, loss] vector per optimiser step (the reference's only collective: MpiAdam, /root/reference
src/distilation/backup/student_rollout.py:658-659,709).  Env shards never communicate."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(total_envs, rank, world):
    """Contiguous global env id range [lo, hi) of `rank`; remainders go to the lowest ranks."""
    base, rem = divmod(int(total_envs), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_gradloss(gradloss, group=None):
    """Sum [flat grad, loss] over ranks in place (KL is a sum, so this equals the gradient on the concatenated batch)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(gradloss, op=dist.ReduceOp.SUM, group=group)
    return gradloss


def max_over_ranks(value, device=None):
    """Max of a python float over ranks (timing)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def rank_checkpoint_path(path, rank, world):
    """Loop checkpoints hold per-shard state (env state, Philox episode counters, `prev` carry): one file per rank when world > 1."""
    return path if world <= 1 else "%s.rank%d" % (path, rank)


def all_ranks_agree(flag, device=None):
    """True only if `flag` is true on every rank (e.g. every rank found its shard's checkpoint: either all resume or none does)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return bool(flag)
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item())
