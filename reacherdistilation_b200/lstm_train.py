"""LSTM distillation loop -- drop-in for /root/reference src/distilation/lstm_train.py:18-201 (`train(train, restore)`), N envs wide.

phase A (:115-137)  the teacher acts, records go to the Dataset, flush on done, until more than 2 * LSTM_BATCH_SIZE episodes exist;
phase B (:141-201)  per env step: one optimiser step per Dataset.training_batches() batch (windows [T,B,.], zero initial state, ob
                    dropout keep_prob, teacher-forced prev_pdflat, KL loss, Adam lr 1e-3); teacher label for the current observation;
                    student action = last row of the LSTM run over Dataset.test_batch(ob) from the CARRIED state (the reference feeds
                    the previous final state back in, :171-182); record ('s'); step; flush on done.
All tensor work is on the device (tcgen05 GEMMs); the checkpoint is a torch.save of the flat parameters + Adam moments (:199).
`thread_state=True` is the loop of backup/lstm_bbpt.py:120-137: the optimiser steps of one env step run over `training_epochs` successive
batches and the final LSTM state of each batch is fed to the next as its initial state (forward value only -- the reference feeds it through
a placeholder, so no gradient crosses the batch boundary); it starts from zeros at every env step."""
import os

import torch

from .config import KEEP_PROB, LSTM_BATCH_SIZE, SEED, STEPS_UNROLLED, TOTAL_EPISODES, TRAINING_EPOCHS, base_path
from .dataset import Dataset
from .dist import all_ranks_agree, rank_checkpoint_path
from .env import VecReacher
from .student_nn import StudentLSTM
from .teacher import TeacherAgent


def train(train=True, restore=False, num_envs=64, total_episodes=TOTAL_EPISODES, iterations=None, seed=SEED, device=0, keep_prob=KEEP_PROB,
          batch_size=LSTM_BATCH_SIZE, generations=16, lr=1e-3, checkpoint=None, verbose=True, thread_state=False, training_epochs=TRAINING_EPOCHS,
          teacher_params=None, teacher_ckpt=None, use_graph=True, save_every_episode=True):
    """`restore`: continue from `checkpoint` (default base_path/student_lstm_b200.pt; saved at every episode boundary like lstm_train.py:199).
    The checkpoint holds the whole loop -- student + Adam moments, env state, the Dataset ring and its sampling counter, the carried acting
    state, the current observation / reward -- so a restored run continues bit-identically; a file with only the student (older runs) restores
    the student like the reference does.  Under torchrun every rank owns `num_envs` envs and its own Dataset; the flat [grad | loss] vector is
    summed over ranks before Adam (the reference's only collective, MpiAdam: backup/student_rollout.py:658-659,709)."""
    from ._lib import MODE_TC
    from .mlp_train import _teacher_for_run
    T = STEPS_UNROLLED
    rank, world = 0, 1
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        rank, world = torch.distributed.get_rank(), torch.distributed.get_world_size()
    env = VecReacher(num_envs=num_envs, seed=seed, device=device, env_offset=rank * num_envs)
    tparams, tdesc = _teacher_for_run(teacher_params, teacher_ckpt, verbose, rank)
    teacher = TeacherAgent(env, params=tparams, mode=MODE_TC)
    student = StudentLSTM(seed=1, device=env.device, lr=lr, max_batch=max(batch_size, num_envs))
    ckpt = checkpoint or os.path.join(base_path, "student_lstm_b200.pt")
    my_ckpt = rank_checkpoint_path(ckpt, rank, world)
    dataset = Dataset(num_envs=num_envs, generations=generations, device=device, seed=seed + 7919 * rank)
    ob = env.reset()
    reward = torch.zeros(num_envs, device=env.device)
    state = student.zero_state(num_envs)                                   # curr_state_batch (:93-94)
    losses, rewards, it, resumed = [], [], 0, False
    if restore:
        sd = torch.load(my_ckpt) if os.path.exists(my_ckpt) else (torch.load(ckpt) if os.path.exists(ckpt) else None)
        full = sd is not None and "dataset" in sd and int(sd.get("num_envs", -1)) == int(num_envs) and int(sd.get("env_offset", -1)) == rank * num_envs
        if all_ranks_agree(full, env.device if world > 1 and torch.distributed.get_backend() == "nccl" else None):
            student.load_state_dict(sd["student"])
            e = sd["env"]
            env.set_state(qpos=e["qpos"], qvel=e["qvel"], target=e["target"], fingertip=e["fingertip"], step=e["step"], episode=e["episode"],
                          qpos_lo=e.get("qpos_lo"))
            dataset.load_state_dict(sd["dataset"])
            ob = env.observe().clone()
            reward, state = sd["reward"].to(env.device), sd["state"].to(env.device)
            it, losses, rewards, resumed = int(sd["it"]), list(sd["losses"]), list(sd["rewards"]), True
        elif sd is not None:
            student.load_state_dict(sd["student"] if "student" in sd else sd)
    if world > 1:
        torch.distributed.broadcast(student.params, src=0)                 # MpiAdam.sync()
    if not train:
        return dict(env=env, teacher=teacher, student=student, dataset=dataset)

    def save():
        os.makedirs(os.path.dirname(my_ckpt) or ".", exist_ok=True)
        torch.save(dict(student=student.state_dict(), env={k: v.cpu() for k, v in env.get_state().items()}, dataset=dataset.state_dict(),
                        reward=reward.cpu(), state=state.cpu(), it=it, losses=losses, rewards=rewards, num_envs=num_envs, env_offset=rank * num_envs,
                        seed=seed), my_ckpt)

    if verbose and rank == 0:
        print("Begin Training! First Accumulate observation with teacher")
    while not resumed and dataset.num_episodes() <= 2 * LSTM_BATCH_SIZE:
        t_pdflat = teacher.pdflat(ob)
        dataset.write(ob, reward, t_pdflat, None, "t")
        ob, reward, new, _ = env.step(t_pdflat[:, :2].contiguous())
        if dataset.last_step() + 1 == 50:
            dataset.flush()
    if verbose and rank == 0:
        print("Accumulated sufficient data points from teacher. now train")
    max_it = iterations if iterations is not None else 50 * max(1, -(-total_episodes // (num_envs * world)))
    # one optimiser step = ONE CUDA-graph launch (rb_lstm_step: forward, KL, BPTT and Adam; the dropout iteration and the Adam step come from a
    # device-side clock) when nothing has to happen between the gradient and the update; static batch buffers keep the captured graph valid
    graph = bool(use_graph) and world == 1 and not thread_state
    dev = env.device
    bufs = (torch.empty((T, batch_size, 11), device=dev), torch.empty((T, batch_size, 4), device=dev), torch.empty((T, batch_size, 4), device=dev),
            torch.empty((T, batch_size, 1), device=dev))
    total_loss = 0.0
    while it < max_it:
        s_thread = student.zero_state(batch_size) if thread_state else None     # lstm_bbpt.py:122 `s = zero_state_batch`
        total_loss = 0.0
        for _ in range(training_epochs):
            ob_b, t_b, prev_b, _prew = dataset.training_batch(batch_size, T, out=bufs)
            if graph:
                student.step(ob_b, prev_b, t_b, None, keep_prob=keep_prob, seed=seed)
            else:
                fin = torch.empty_like(s_thread) if thread_state else None
                student.loss_grad(ob_b, prev_b, t_b, s_thread, keep_prob=keep_prob, seed=seed, iteration=student.t, final_state_out=fin)
                if world > 1:
                    torch.distributed.all_reduce(student.gradloss)
                student.adam_step()
                s_thread = fin
            total_loss = total_loss + student.gradloss[student.P]               # `total_loss += l` (lstm_train.py:161, lstm_bbpt.py:137)
        t_pdflat = teacher.pdflat(ob)
        tb = dataset.test_batch(ob, steps=T)
        ob_w, prev_w = (tb[0], tb[1]) if num_envs > 1 else (tb[0][:, -1:, :].contiguous(), tb[1][:, -1:, :].contiguous())
        s_win, state = student.forward(ob_w, prev_w, state)
        s_pdflat = s_win[T - 1]
        dataset.write(ob, reward, t_pdflat, s_pdflat, "s")
        ob, reward, new, _ = env.step(s_pdflat[:, :2].contiguous())
        it += 1
        if dataset.last_step() + 1 == 50:
            dataset.flush()
            losses.append(float(total_loss)); rewards.append(float(reward.mean()))
            if verbose and rank == 0:
                print("************** Episode %d ****************" % dataset.num_episodes())
                print("recent loss: %f " % losses[-1])
            if save_every_episode:
                save()                                                          # lstm_train.py:199 saves the student every episode
    save()
    return dict(losses=losses, rewards=rewards, iterations=it, env=env, teacher=teacher, student=student, dataset=dataset, checkpoint=my_ckpt,
                teacher_desc=tdesc, resumed=resumed)
