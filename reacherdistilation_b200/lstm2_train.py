"""Two-headed LSTM distillation loop -- drop-in for /root/reference src/distilation/backup/student_rollout.py:260-588
(`lstm_train(train, keep, lstm_trained_data_path, restore)`), N lock-step envs wide.

history (:345, the reference's python lists, here device tensors [time, N, .]):  ob, teacher pdflat (mean | logstd; std = exp(logstd) is
        recomputed where needed), stepped action, reward
phase A (:467-488)  the teacher acts with its mean until at least one episode (EPISODE_STEPS records) exists
phase B (:494-580)  per env step: teacher label for the current observation; ONE optimiser step on a batch of random windows
                    (generate_training_set :205-240: a completed episode, a start 1 .. EPISODE_STEPS - T - 1 inside it; ob / reward / teacher rows
                    [index, index + T), stepped actions [index - 1, index - 1 + T)), zero initial state, ob dropout `keep`, loss = KL + squared
                    reward error, Adam 1e-3; then the student acts: forward (keep_prob 1) over each env's last T (ob, stepped action) rows from the
                    CARRIED state (generate_test_set :243-258 puts that one window in the last batch slot; here every env is a slot), action =
                    mean of the last unrolled step (s_ac :318); step; gamma-discounted return per episode (:540); checkpoint every episode (:556).
The window gather is torch indexing on the device tensors (data movement only); all arithmetic runs in rb_lstm2_* / rb_adam_step."""
import os

import numpy as np
import torch

from .config import SEED, base_path
from .env import VecReacher
from .student_nn import StudentLSTM2, lstm2_spec
from .teacher import TeacherAgent

EPISODE_STEPS = 50          # backup/student_rollout.py:45
GAMMA = 0.99                # :46
TOTAL_EPISODES = 2          # :35


def lstm_train(train=True, keep=0.5, lstm_trained_data_path=None, restore=False, num_envs=64, batch_size=100, units=100, steps=20, carry_state=False,
               reward_hidden=(64,), total_episodes=TOTAL_EPISODES, iterations=None, seed=SEED, device=0, lr=1e-3, teacher_params=None, teacher_ckpt=None,
               verbose=True):
    """`units` / `steps` / `batch_size`: NUM_UNITS / STEPS_UNROLLED / LSTM_BATCH_SIZE (:38-50 ships 1 / 2 / 2 for debugging next to 100 / 20 / 100).
    Returns dict(losses, rets, ...) like the arrays the reference saves at every episode boundary (:556-565)."""
    from ._lib import MODE_TC
    from .mlp_train import _teacher_for_run
    T, N = int(steps), int(num_envs)
    assert EPISODE_STEPS - T - 1 >= 1, "windows are drawn from inside one 50-step episode (:219)"
    env = VecReacher(num_envs=N, seed=seed, device=device)
    dev = env.device
    tparams, tdesc = _teacher_for_run(teacher_params, teacher_ckpt, verbose, 0)
    teacher = TeacherAgent(env, params=tparams, mode=MODE_TC)
    student = StudentLSTM2(spec=lstm2_spec(units, T, carry_state, reward_hidden=tuple(reward_hidden)), seed=1, device=dev, lr=lr)
    ckpt = lstm_trained_data_path or os.path.join(base_path, "student_lstm2_b200.pt")
    cap = EPISODE_STEPS * (int(total_episodes) + 4) + 8                    # history rows: phase A (one episode) + one row per env step
    if iterations is not None:
        cap = max(cap, 2 * EPISODE_STEPS + int(iterations) + 8)
    h_ob, h_t = torch.zeros((cap, N, 11), device=dev), torch.zeros((cap, N, 4), device=dev)
    h_ac, h_rew = torch.zeros((cap, N, 2), device=dev), torch.zeros((cap, N), device=dev)
    n_ob = n_act = 0                       # len(ob_list), len(stepped_action_list) == len(reward_list)
    episodes, losses, rets = 0, [], []
    gen = torch.Generator(device="cpu").manual_seed(int(seed))
    if restore and os.path.exists(ckpt):
        sd = torch.load(ckpt)
        student.load_state_dict(sd["student"])
        if train and "h_ob" in sd and sd["h_ob"].shape[1] == N:      # the reference reloads its lists too (:423-461)
            n_ob, n_act, episodes, losses, rets = int(sd["n_ob"]), int(sd["n_act"]), int(sd["episodes"]), list(sd["losses"]), list(sd["rets"])
            if n_ob + (int(iterations) if iterations is not None else EPISODE_STEPS * (int(total_episodes) + 1)) + 8 > cap:      # room for the continued run
                grow = n_ob + (int(iterations) if iterations is not None else EPISODE_STEPS * (int(total_episodes) + 1)) + 8
                h_ob, h_t = torch.zeros((grow, N, 11), device=dev), torch.zeros((grow, N, 4), device=dev)
                h_ac, h_rew = torch.zeros((grow, N, 2), device=dev), torch.zeros((grow, N), device=dev)
                cap = grow
            for dst, key in ((h_ob, "h_ob"), (h_t, "h_t"), (h_ac, "h_ac"), (h_rew, "h_rew")):
                dst[:sd[key].shape[0]].copy_(sd[key])
            gen.set_state(sd["gen"])
    elif restore and verbose:
        print("attempt to restore trained data but %s does not exist" % ckpt)
    if not train:
        return dict(env=env, teacher=teacher, student=student)

    def save():
        os.makedirs(os.path.dirname(ckpt) or ".", exist_ok=True)
        torch.save(dict(student=student.state_dict(), h_ob=h_ob[:n_ob].cpu(), h_t=h_t[:n_ob].cpu(), h_ac=h_ac[:n_act].cpu(), h_rew=h_rew[:n_act].cpu(),
                        n_ob=n_ob, n_act=n_act, episodes=episodes, losses=losses, rets=rets, gen=gen.get_state()), ckpt)

    ob = env.reset()
    h_ob[n_ob].copy_(ob); n_ob += 1
    if verbose:
        print("Begin Training! First Accumulate observation with teacher")
    while n_ob < EPISODE_STEPS + 1 and episodes == 0:                       # phase A
        t_pd = teacher.pdflat(ob)
        h_t[n_ob - 1].copy_(t_pd)
        act = t_pd[:, :2].contiguous()
        ob, reward, new, _ = env.step(act)
        h_ac[n_act].copy_(act); h_rew[n_act].copy_(reward); n_act += 1
        if n_act % EPISODE_STEPS == 0:                                         # `new`: lock-step TimeLimit, the env has reset itself
            episodes += 1
        h_ob[n_ob].copy_(ob); n_ob += 1
    if verbose:
        print("Accumulated %d data points from teacher. now train" % n_act)
    max_it = iterations if iterations is not None else EPISODE_STEPS * (int(total_episodes) + 1)
    state = student.zero_state(N)
    ret = torch.zeros(N, device=dev)
    timestep, it, loss = 0, 0, None
    ar = torch.arange(T, device=dev)
    # static window buffers: one optimiser step = ONE CUDA-graph launch (rb_lstm2_step) whose captured pointers stay valid from step to step
    bufs = (torch.empty((T, batch_size, 11), device=dev), torch.empty((T, batch_size, 2), device=dev), torch.empty((T, batch_size, 4), device=dev),
            torch.empty((T, batch_size), device=dev))
    while it < max_it and n_ob < cap - 1:
        t_pd = teacher.pdflat(ob)
        h_t[n_ob - 1].copy_(t_pd)
        # ---- generate_training_set (:205-240): batch_size random windows out of the completed episodes ----------------------------------------
        ep = torch.randint(0, max(episodes, 1), (batch_size,), generator=gen)
        start = torch.randint(1, EPISODE_STEPS - T, (batch_size,), generator=gen)      # randint(1, EPISODE_STEPS - T - 1) inclusive
        e_ix = torch.randint(0, N, (batch_size,), generator=gen).to(dev)
        index = (ep * EPISODE_STEPS + start).to(dev)
        rows = index[None, :] + ar[:, None]                                              # [T, B]
        bufs[0].copy_(h_ob[rows, e_ix[None, :]]); bufs[2].copy_(h_t[rows, e_ix[None, :]]); bufs[3].copy_(h_rew[rows, e_ix[None, :]])
        bufs[1].copy_(h_ac[rows - 1, e_ix[None, :]])
        student.step(bufs[0], bufs[1], bufs[2], bufs[3], None, keep_prob=keep, seed=seed, sample_id0=0)
        loss = student.gradloss[student.P]
        # ---- generate_test_set (:243-258) + acting (:527-535): every env's last T rows, from the carried state -------------------------------
        lo = n_ob - T
        ob_t = h_ob[lo:n_ob] if lo >= 0 else torch.cat([torch.zeros((-lo, N, 11), device=dev), h_ob[:n_ob]])
        la = n_act - T
        ac_t = h_ac[la:n_act] if la >= 0 else torch.cat([torch.zeros((-la, N, 2), device=dev), h_ac[:n_act]])
        s_win, _, state = student.forward(ob_t, ac_t, state)
        s_action = s_win[T - 1, :, :2].contiguous()
        ob, reward, new, _ = env.step(s_action)
        ret = ret + (GAMMA ** timestep) * reward
        timestep += 1
        h_ac[n_act].copy_(s_action); h_rew[n_act].copy_(reward); n_act += 1
        it += 1
        if n_act % EPISODE_STEPS == 0:
            losses.append(float(loss)); rets.append(float(ret.mean()))
            if verbose:
                print("************** Episode %d ****************" % episodes)
                print("Total loss: %f   actual return: %f" % (losses[-1], rets[-1]))
            timestep = 0
            ret = torch.zeros(N, device=dev)
            episodes += 1
        h_ob[n_ob].copy_(ob); n_ob += 1
        if n_act % EPISODE_STEPS == 0:
            save()
        if iterations is None and episodes > total_episodes:
            break
    save()
    return dict(losses=losses, rets=rets, iterations=it, episodes=episodes, env=env, teacher=teacher, student=student, checkpoint=ckpt,
                teacher_desc=tdesc, last_loss=None if loss is None else float(loss))
