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
from .env import VecReacher
from .student_nn import StudentLSTM
from .teacher import TeacherAgent


def train(train=True, restore=False, num_envs=64, total_episodes=TOTAL_EPISODES, iterations=None, seed=SEED, device=0, keep_prob=KEEP_PROB,
          batch_size=LSTM_BATCH_SIZE, generations=16, lr=1e-3, checkpoint=None, verbose=True, thread_state=False, training_epochs=TRAINING_EPOCHS):
    from ._lib import MODE_TC
    T = STEPS_UNROLLED
    env = VecReacher(num_envs=num_envs, seed=seed, device=device)
    teacher = TeacherAgent(env, restore=restore, mode=MODE_TC)
    student = StudentLSTM(seed=1, device=env.device, lr=lr, max_batch=max(batch_size, num_envs))
    ckpt = checkpoint or os.path.join(base_path, "student_lstm_b200.pt")
    if restore and os.path.exists(ckpt):
        student.load_state_dict(torch.load(ckpt))
    dataset = Dataset(num_envs=num_envs, generations=generations, device=device, seed=seed)
    ob = env.reset()
    reward = torch.zeros(num_envs, device=env.device)
    if not train:
        return dict(env=env, teacher=teacher, student=student, dataset=dataset)
    if verbose:
        print("Begin Training! First Accumulate observation with teacher")
    while dataset.num_episodes() <= 2 * LSTM_BATCH_SIZE:
        t_pdflat = teacher.pdflat(ob)
        dataset.write(ob, reward, t_pdflat, None, "t")
        ob, reward, new, _ = env.step(t_pdflat[:, :2].contiguous())
        if dataset.last_step() + 1 == 50:
            dataset.flush()
    if verbose:
        print("Accumulated sufficient data points from teacher. now train")
    state = student.zero_state(num_envs)                                   # curr_state_batch (:93-94)
    losses, rewards, it = [], [], 0
    max_it = iterations if iterations is not None else 50 * max(1, -(-total_episodes // num_envs))
    total_loss = 0.0
    while it < max_it:
        s_thread = student.zero_state(batch_size) if thread_state else None     # lstm_bbpt.py:122 `s = zero_state_batch`
        total_loss = 0.0
        for _ in range(training_epochs):
            ob_b, t_b, prev_b, _prew = dataset.training_batch(batch_size, T)
            fin = torch.empty_like(s_thread) if thread_state else None
            student.loss_grad(ob_b, prev_b, t_b, s_thread, keep_prob=keep_prob, seed=seed, iteration=student.t, final_state_out=fin)
            student.adam_step()
            total_loss = total_loss + student.gradloss[student.P]               # `total_loss += l` (lstm_train.py:161, lstm_bbpt.py:137)
            s_thread = fin
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
            if verbose:
                print("************** Episode %d ****************" % dataset.num_episodes())
                print("recent loss: %f " % losses[-1])
            if checkpoint is not None:
                os.makedirs(os.path.dirname(ckpt) or ".", exist_ok=True)
                torch.save(student.state_dict(), ckpt)
    return dict(losses=losses, rewards=rewards, iterations=it, env=env, teacher=teacher, student=student, dataset=dataset)
