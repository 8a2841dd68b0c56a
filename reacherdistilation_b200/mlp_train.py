"""MLP distillation loop -- drop-in for /root/reference src/distilation/mlp_train.py:18-204 (`train(train, restore)`),
and the batched variant backup/student_rollout.py:618-740 (`train_student`).

The reference runs ONE env and one 200-sample optimiser step per env step, through three sess.run calls and Python list
packing.  Here N lock-step envs live on the GPU; one DAgger iteration is
    observe   ob_k, teacher label t_k = teacher(ob_k), student input x_k            (rb_dagger_observe)
    loss/grad s_k = student(x_k); L = KL(s_k || t_k) summed; flat gradient          (rb_student_loss_grad)
    exchange  one sum all-reduce of [grad, loss] across ranks (MpiAdam.update)      (NCCL, only if world_size > 1)
    update    TF-form Adam                                                           (rb_adam_step)
    act       env.step(mean(s_k)) -> reward, done; prev-pdflat / prev-reward shift   (rb_dagger_act)
with no host round trip; the loss stays on the device until asked for.
"""
import os

import numpy as np
import torch

from . import _lib
from ._lib import LOSS_KL_ST, MODE_FP32, STUDENT_MLP, check, lib, ptr, stream_ptr
from .config import KEEP_PROB, MLP_BATCH_SIZE, NUM_ENVS, SEED, base_path
from .dist import all_ranks_agree, rank_checkpoint_path
from .env import VecReacher
from .student_nn import StudentNet
from .teacher import TeacherAgent, load_teacher_params


class DaggerTrainer:
    def __init__(self, num_envs=NUM_ENVS, seed=SEED, device=0, student_kind=STUDENT_MLP, keep_prob=KEEP_PROB, mode=MODE_FP32,
                 teacher_params=None, teacher_seed=0, student_seed=1, env_offset=0, loss_kind=LOSS_KL_ST, lr=None, eps=None,
                 process_group=None, average_grads=False, student_params=None, student_mode=None, fused_allreduce=None, use_graph=None, solo=False):
        import ctypes as C
        self.env = VecReacher(num_envs=num_envs, seed=seed, device=device, env_offset=env_offset)
        self.device = self.env.device
        self.teacher = TeacherAgent(self.env, params=teacher_params, seed=teacher_seed, mode=mode)
        if student_mode is None:   # tensor-core student kernels only where this build has them; never a silent substitution of `mode`
            student_mode = mode if lib().rb_student_mode_available(mode) else MODE_FP32
        self.student_mode = student_mode
        self.student = StudentNet(kind=student_kind, seed=student_seed, device=self.device, mode=student_mode, lr=lr, eps=eps, params=student_params)
        self.mode, self.loss_kind, self.n = mode, loss_kind, int(num_envs)
        h = C.c_void_p()
        check(lib().rb_dagger_create(C.byref(h), self.env._h, student_kind, float(keep_prob)))
        self._h = h
        n, dev = self.n, self.device
        with torch.cuda.device(dev):
            self.obs = torch.empty((n, 11), device=dev)
            self.t_pd = torch.empty((n, 4), device=dev)
            self.s_pd = torch.empty((n, 4), device=dev)
            self.x = torch.empty((n, self.student.in_dim), device=dev)
            # The reference drops observations only in the TRAINING batch (keep_prob = KEEP_PROB, mlp_train.py:151) and acts on the clean
            # observation (keep_prob 1, mlp_train.py:171-186): with dropout on, the observe kernel also writes the un-dropped input rows and the
            # student acts on those (s_pd); the loss and the gradient use the dropped rows (x).  The 2x64 student has no input dropout.
            self.x_act = torch.empty((n, self.student.in_dim), device=dev) if (student_kind == STUDENT_MLP and keep_prob < 1.0) else None
            self.s_train = torch.empty((n, 4), device=dev) if self.x_act is not None else None
            self.rew = torch.empty((n,), device=dev)
            self.done = torch.empty((n,), dtype=torch.uint8, device=dev)
        self.pg = process_group
        self.world = 1
        if not solo and (process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized())):
            self.world = torch.distributed.get_world_size(process_group)      # solo: a single-rank trainer inside a multi-rank job (no exchange)
        self.grad_scale = (1.0 / self.world) if average_grads else 1.0   # MpiAdam averages (backup :709); KL sum = concatenated batch
        # world > 1: exchange the gradient inside the student kernel over NVLink peer memory (tensor-core path) unless told to use NCCL
        self.fused_allreduce = (self.world > 1 and self.student_mode == _lib.MODE_TC) if fused_allreduce is None else bool(fused_allreduce)
        if self.fused_allreduce:
            self.student.enable_peer_exchange(process_group)
        # one CUDA-graph launch per iteration (device-side step clock) whenever the whole iteration is on the tensor-core path
        graph_ok = self.mode == _lib.MODE_TC and self.student_mode == _lib.MODE_TC and (self.world == 1 or self.fused_allreduce)
        self.use_graph = graph_ok if use_graph is None else (bool(use_graph) and graph_ok)
        self._clock_synced = False
        self.iteration = 0
        self.env.reset()

    def sync_params(self):
        """MpiAdam.sync() (backup/student_rollout.py:659): broadcast rank-0 student parameters."""
        if self.world > 1:
            torch.distributed.broadcast(self.student.params, src=0, group=self.pg)

    def step(self):
        """One DAgger iteration over all envs.  Asynchronous; returns nothing (loss: self.last_loss())."""
        L, st = lib(), stream_ptr()
        if self.use_graph:
            return self._step_graph(L, st)
        self._clock_synced = False
        check(L.rb_dagger_observe(self._h, ptr(self.teacher.params), self.iteration, ptr(self.obs), ptr(self.t_pd), ptr(self.x), ptr(self.x_act),
                                  self.mode, st))
        s_train = self.s_pd
        if self.x_act is not None:         # acting forward on the un-dropped rows, with the parameters this iteration's gradient is taken at
            self.student.forward(self.x_act, out=self.s_pd)
            s_train = self.s_train
        if self.fused_allreduce:
            self.student.step_dp(self.x, self.t_pd, self.loss_kind, s_out=s_train, grad_scale=self.grad_scale)
        elif self.world > 1:
            self.student.loss_grad(self.x, self.t_pd, self.loss_kind, s_out=s_train)
            torch.distributed.all_reduce(self.student.gradloss, group=self.pg)
            self.student.adam_step(self.grad_scale)
        else:
            self.student.step(self.x, self.t_pd, self.loss_kind, s_out=s_train, grad_scale=self.grad_scale)
        check(L.rb_dagger_act(self._h, ptr(self.s_pd), ptr(self.t_pd), ptr(self.rew), ptr(self.done), st))
        self.iteration += 1

    def _step_graph(self, L, st):
        """rb_dagger_step: the iteration as one captured CUDA graph; per-step values come from the device-side clock."""
        stu = self.student
        if not self._clock_synced:         # (re)load the device clock from the host counters after any non-graph step
            check(L.rb_dagger_set_clock(self._h, self.iteration, stu.t, getattr(stu, "_px_epoch", 0), st))
            self._clock_synced = True
        if self.fused_allreduce:
            se, so, fl = stu._px_slots[0].ctypes.data, stu._px_slots[1].ctypes.data, stu._px_flags.ctypes.data
            rank, world = stu._px_rank, stu._px_world
        else:
            se = so = fl = None
            rank, world = 0, 1
        check(L.rb_dagger_step(self._h, ptr(self.teacher.params), ptr(stu.params), ptr(stu.m), ptr(stu.v), ptr(stu.gradloss), ptr(stu.workspace),
                               ptr(self.obs), ptr(self.t_pd), ptr(self.x), ptr(self.x_act), ptr(self.s_pd), ptr(self.rew), ptr(self.done), self.loss_kind,
                               stu.lr, stu.beta1, stu.beta2, stu.eps, self.grad_scale, rank, world, se, so, fl, 1, st))
        self.iteration += 1
        stu.t += 1
        if self.fused_allreduce:
            stu._px_epoch += 1

    def last_loss(self):
        return self.student.gradloss[self.student.P]

    def wait_loss(self):
        """Loss of the iteration the last step() issued, as a host float.  On the CUDA-graph path the last kernel of the iteration posts it
        into page-locked host memory and this call polls for it (rb_dagger_wait_loss: no stream synchronise, no copy); otherwise it is a
        plain device read."""
        if self.use_graph:
            import ctypes as C
            out = C.c_float()
            check(lib().rb_dagger_wait_loss(self._h, self.iteration, C.byref(out)))
            return out.value
        return float(self.last_loss())

    def state_dict(self):
        """Everything an exact resume needs (the reference checkpoints the student only, lstm_train.py:102-107,199; here also the loop state, so
        that a restored run continues bit-identically): student parameters + Adam moments, env state (qpos, qvel, target, stale fingertip, step,
        Philox episode counter), the handle's `prev` / `prew` carry and the iteration counter."""
        n, dev = self.n, self.device
        prev_t, prr, lr_ = torch.empty((n, 4), device=dev), torch.empty((n,), device=dev), torch.empty((n,), device=dev)
        check(lib().rb_dagger_get_state(self._h, ptr(prev_t), ptr(prr), ptr(lr_), stream_ptr()))
        env = {k: v.cpu() for k, v in self.env.get_state().items()}
        return dict(student=self.student.state_dict(), env=env, prev_t_pdflat=prev_t.cpu(), prev_rec_rew=prr.cpu(), last_reward=lr_.cpu(),
                    iteration=self.iteration, num_envs=n, seed=self.env.seed, env_offset=self.env.env_offset)

    def load_state_dict(self, sd):
        if int(sd["num_envs"]) != self.n or int(sd["seed"]) != int(self.env.seed) or int(sd["env_offset"]) != int(self.env.env_offset):
            raise ValueError("checkpoint is for num_envs=%d seed=%d env_offset=%d" % (sd["num_envs"], sd["seed"], sd["env_offset"]))
        self.student.load_state_dict(sd["student"])
        e = sd["env"]
        self.env.set_state(qpos=e["qpos"], qvel=e["qvel"], target=e["target"], fingertip=e["fingertip"], step=e["step"], episode=e["episode"],
                           qpos_lo=e.get("qpos_lo"))
        dev = self.device
        prev_t, prr, lr_ = (sd[k].to(dev, torch.float32).contiguous() for k in ("prev_t_pdflat", "prev_rec_rew", "last_reward"))
        check(lib().rb_dagger_set_state(self._h, ptr(prev_t), ptr(prr), ptr(lr_), stream_ptr()))
        torch.cuda.current_stream().synchronize()          # the temporaries must outlive the copies
        self.iteration = int(sd["iteration"])
        self._clock_synced = False                         # the device-side step clock is reloaded from the host counters

    def close(self):
        if self._h:
            lib().rb_dagger_destroy(self._h)
            self._h = None
        self.env.close()


def _teacher_for_run(teacher_params, teacher_ckpt, verbose, rank=0):
    """Teacher weights of a training run: supplied vector / checkpoint file / `base_path/teacher.npz`; without any of them the run distils a
    SEEDED UNTRAINED teacher and says so (the reference's teacher.ckpt is not in its repository) -- the returned description goes into the
    result dict so that no run mistakes the synthetic teacher for a trained one."""
    if teacher_params is None and teacher_ckpt is None and os.path.exists(os.path.join(base_path, "teacher.npz")):
        teacher_ckpt = os.path.join(base_path, "teacher.npz")
    p, desc = load_teacher_params(teacher_params, teacher_ckpt)
    if p is None:
        desc = "SYNTHETIC: seeded baselines-initialised (untrained) teacher -- no teacher_params / teacher_ckpt given"
        if verbose and rank == 0:
            print("teacher: " + desc)
    return p, desc


def train(train=True, restore=False, num_envs=NUM_ENVS, iterations=None, total_episodes=5000, seed=SEED, device=0, student_kind=STUDENT_MLP,
          keep_prob=KEEP_PROB, mode=MODE_FP32, warmup_episodes=2 * MLP_BATCH_SIZE + 1, log_every=50, checkpoint=None, verbose=True,
          teacher_params=None, teacher_ckpt=None):
    """Same entry point as the reference (`mlp_train.train(train, restore)`, main.py:26-27).
    teacher_params / teacher_ckpt: the weights `teacher.py:17-20` restores (flat vector, or an .npz of the baselines variables).
    The loop state is saved to `checkpoint` (default base_path/student_mlp_b200.pt, the file `restore` and `main.py -ch` read) at the end.

    Phase A (mlp_train.py:120-139): every env plays `warmup_episodes / num_envs` (>= 1) teacher episodes into the device buffer.
    Phase B (mlp_train.py:143-204): DAgger iterations until `total_episodes` episodes (5000 in the reference) or `iterations`.
    Returns a dict with the loss curve (one summed KL per logged iteration) and mean per-step reward."""
    rank, world = 0, 1
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        rank, world = torch.distributed.get_rank(), torch.distributed.get_world_size()
    tparams, tdesc = _teacher_for_run(teacher_params, teacher_ckpt, verbose, rank)
    tr = DaggerTrainer(num_envs=num_envs, seed=seed, device=device, student_kind=student_kind, keep_prob=keep_prob, mode=mode,
                       env_offset=rank * num_envs, teacher_params=tparams)
    ckpt = checkpoint or os.path.join(base_path, "student_mlp_b200.pt")
    my_ckpt = rank_checkpoint_path(ckpt, rank, world)    # per-shard loop state: one file per rank when world > 1
    resumed = False
    if restore:
        sd = torch.load(my_ckpt) if os.path.exists(my_ckpt) else (torch.load(ckpt) if os.path.exists(ckpt) else None)
        full = (sd is not None and "student" in sd and int(sd.get("num_envs", -1)) == int(num_envs)
                and int(sd.get("env_offset", -1)) == rank * num_envs)
        if all_ranks_agree(full, tr.device if world > 1 and torch.distributed.get_backend() == "nccl" else None):
            tr.load_state_dict(sd)                       # full loop state: continue exactly where the saved run stopped
            resumed = True
        elif sd is not None:                             # student-only checkpoint (or another shard shape): the reference's restore
            tr.student.load_state_dict(sd["student"] if "student" in sd else sd)
    tr.sync_params()
    if not train:
        return dict(trainer=tr)
    if verbose and rank == 0:
        print("Begin Training! First Accumulate observation with teacher")
    teacher_reward = None
    if not resumed:                                      # a resumed run is already past phase A
        eps_per_env = max(1, -(-warmup_episodes // (num_envs * world)))
        warm = tr.env.rollout_policy(tr.teacher.params, 50 * eps_per_env, nout=2, mode=mode)
        teacher_reward = float(warm["rew"].mean())
    if verbose and rank == 0:
        print("Accumulated sufficient data points from teacher. now train")
    if iterations is None:
        iterations = 50 * max(1, -(-total_episodes // (num_envs * world)))
    losses, rewards = [], []
    for it in range(iterations):
        tr.step()
        if (it + 1) % log_every == 0 or it == iterations - 1:
            losses.append(float(tr.last_loss()))
            rewards.append(float(tr.rew.mean()))
            if verbose and rank == 0:
                print("************** Episode %d ****************" % ((it + 1) // 50 * num_envs * world))
                print("recent loss: %f " % losses[-1])
    os.makedirs(os.path.dirname(my_ckpt) or ".", exist_ok=True)       # always: `-r` and `-ch` read this file (the reference saves every episode)
    torch.save(tr.state_dict(), my_ckpt)
    out = dict(losses=losses, rewards=rewards, teacher_reward=teacher_reward, iterations=iterations, trainer=tr, checkpoint=my_ckpt, teacher=tdesc,
               resumed=resumed)
    return out


def train_replay(train=True, restore=False, num_envs=64, total_episodes=5000, iterations=None, seed=SEED, device=0, keep_prob=KEEP_PROB,
                 mode=MODE_FP32, batch_size=MLP_BATCH_SIZE, steps_unrolled=None, generations=64, lr=None, verbose=True, teacher_params=None,
                 teacher_ckpt=None):
    """The reference loop in its own shape (mlp_train.py:110-204), N envs wide, with the device Dataset as replay buffer:

    phase A (:120-139)  teacher acts; every record {ob, reward, t_pdflat, s=0, 't'} is written; flush on done, until more than
                        2 * MLP_BATCH_SIZE episodes exist;
    phase B (:143-204)  per env step: one optimiser step per Dataset.training_batches() batch (windows [T,B,.] drawn from the
                        buffer, student input = [dropout(ob), prev, prew], loss.py KL vs the recorded teacher pdflat); teacher label
                        for the current observation; student action from [ob, prev, prew] with keep_prob 1 (the last row of
                        Dataset.test_batch -- the MLP graph only uses that row, :63-66); write the record ('s'); step with the
                        student's mean; flush on done.
    The reference feeds random noise for prev/prew in this file ("TODO: revert this back", :157-158); the reverted form is used."""
    from .config import STEPS_UNROLLED
    from .dataset import Dataset
    from .student_nn import student_mlp_input
    T = steps_unrolled or STEPS_UNROLLED
    env = VecReacher(num_envs=num_envs, seed=seed, device=device)
    teacher = TeacherAgent(env, params=_teacher_for_run(teacher_params, teacher_ckpt, verbose)[0], mode=mode)
    student = StudentNet(kind=STUDENT_MLP, seed=1, device=env.device, mode=mode, lr=lr)
    dataset = Dataset(num_envs=num_envs, generations=generations, device=device, seed=seed)
    ob = env.reset()
    reward = torch.zeros(num_envs, device=env.device)
    if not train:
        return dict(env=env, teacher=teacher, student=student, dataset=dataset)
    if verbose:
        print("Begin Training! First Accumulate observation with teacher")
    while dataset.num_episodes() <= 2 * MLP_BATCH_SIZE:
        t_pdflat = teacher.pdflat(ob)
        dataset.write(ob, reward, t_pdflat, None, "t")
        ob, reward, new, _ = env.step(t_pdflat[:, :2].contiguous())
        if dataset.last_step() + 1 == 50:                      # lock-step envs: every episode ends on the same step (TimeLimit 50)
            dataset.flush()
    if verbose:
        print("Accumulated sufficient data points from teacher. now train")
    losses, rewards, it = [], [], 0
    max_it = iterations if iterations is not None else 50 * max(1, -(-total_episodes // num_envs))
    while it < max_it:
        total_loss = 0.0
        for (ob_b, t_b, prev_b, prew_b) in dataset.training_batches(batch_size, T):
            x = student_mlp_input(ob_b.reshape(-1, 11), prev_b.reshape(-1, 4), prew_b.reshape(-1), keep_prob, seed, 0, it)
            student.step(x, t_b.reshape(-1, 4))
            total_loss = student.gradloss[student.P]
        t_pdflat = teacher.pdflat(ob)
        _, prev_p, prev_r = dataset.test_batch(ob, steps=1) if num_envs > 1 else [a[:, -1:, :] for a in dataset.test_batch(ob, steps=1)]
        x_act = student_mlp_input(ob, prev_p.reshape(-1, 4), prev_r.reshape(-1), 1.0, seed, 0, it)
        s_pdflat = student.forward(x_act)
        dataset.write(ob, reward, t_pdflat, s_pdflat, "s")
        ob, reward, new, _ = env.step(s_pdflat[:, :2].contiguous())
        it += 1
        if dataset.last_step() + 1 == 50:
            dataset.flush()
            losses.append(float(total_loss)); rewards.append(float(reward.mean()))
            if verbose:
                print("************** Episode %d ****************" % dataset.num_episodes())
                print("recent loss: %f " % losses[-1])
    return dict(losses=losses, rewards=rewards, iterations=it, env=env, teacher=teacher, student=student, dataset=dataset)
