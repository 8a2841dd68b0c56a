"""TEST INFRASTRUCTURE (oracle): numpy restatement of the reference Dataset semantics, /root/reference
src/distilation/dataset.py:118-290, for N lock-step envs.  Episodes are Python lists of dict records exactly as in the reference
(keys ob, rew, t, s, with, prev, prew); sampling uses Philox4x32-10 keyed (seed; draw, b, 0, stream 3) with multiply-shift
range reduction (Python's `random` stream is unpinned in the reference).  Only tests may import this module."""
import numpy as np

from .philox_np import philox4x32_10

EPISODE_STEPS, STEPS_UNROLLED, LSTM_BATCH_SIZE = 50, 10, 20
STREAM_DATASET = 3


def _u32(seed, c0, c1):
    return int(philox4x32_10(seed, np.array([c0]), np.array([c1]), np.array([0]), np.array([STREAM_DATASET]))[0][0])


class DatasetOracle:
    def __init__(self, num_envs=1, generations=64, seed=0):
        self.n, self.G, self.seed = num_envs, generations, seed
        self.data_in_memory = []                       # complete episodes, oldest first (ring of G generations x n envs)
        self.curr = [[] for _ in range(num_envs)]      # curr_episode per env
        self.num_total_episodes = 0

    # dataset.py:151-164
    @staticmethod
    def pdflat_at(episode, k):
        return np.zeros(4) if k < 0 else episode[k]["t"]

    @staticmethod
    def rew_at(episode, k):
        return 0.0 if k < 0 else episode[k]["rew"]

    def write(self, ob, reward=None, t_pdflat=None, s_pdflat=None, stepped_with="t"):          # dataset.py:118-143
        for e in range(self.n):
            ep = self.curr[e]
            step = dict(ob=np.asarray(ob[e], np.float64), rew=0.0 if reward is None else float(reward[e]),
                        t=np.zeros(4) if t_pdflat is None else np.asarray(t_pdflat[e], np.float64),
                        s=np.zeros(4) if s_pdflat is None else np.asarray(s_pdflat[e], np.float64))
            step["with"] = stepped_with
            step["prev"] = self.pdflat_at(ep, len(ep) - 1)
            step["prew"] = self.rew_at(ep, len(ep) - 1)
            ep.append(step)
        if len(self.data_in_memory) == self.G * self.n and len(self.curr[0]) == 1:
            del self.data_in_memory[:self.n]            # the ring starts overwriting its oldest generation

    def flush(self):                                                                            # dataset.py:146-149
        assert all(len(ep) == EPISODE_STEPS for ep in self.curr)
        self.data_in_memory.extend(self.curr)
        self.curr = [[] for _ in range(self.n)]
        self.num_total_episodes += self.n

    def training_batch(self, draw, B=LSTM_BATCH_SIZE, T=STEPS_UNROLLED):                        # dataset.py:184-210
        navail = len(self.data_in_memory)
        start = (_u32(self.seed, draw, 0xFFFFFFFF) * (EPISODE_STEPS - T + 1)) >> 32
        eps = [(_u32(self.seed, draw, b) * navail) >> 32 for b in range(B)]
        ser = lambda key: np.transpose(np.array([[np.atleast_1d(self.data_in_memory[e][k][key]) for k in range(start, start + T)] for e in eps]),
                                       (1, 0, 2))
        return ser("ob"), ser("t"), ser("prev"), ser("prew"), np.array(eps), start

    def test_batch(self, ob, T=STEPS_UNROLLED):                                                 # dataset.py:213-290 (intended semantics,
        obs, prevs, prews = [], [], []                                                         # tests/dataset_unit_test.py:13-94)
        for e in range(self.n):
            ep = self.curr[e]
            start = len(ep) - T + 1
            o = [np.zeros(11)] * max(0, -start) + [r["ob"] for r in ep[max(0, start):]] + [np.asarray(ob[e], np.float64)]
            p = [np.zeros(4)] * max(0, -start) + [r["prev"] for r in ep[max(0, start):]] + [self.pdflat_at(ep, len(ep) - 1)]
            w = [0.0] * max(0, -start) + [r["prew"] for r in ep[max(0, start):]] + [self.rew_at(ep, len(ep) - 1)]
            obs.append(np.array(o)); prevs.append(np.array(p)); prews.append(np.array(w)[:, None])
        return np.stack(obs, 1), np.stack(prevs, 1), np.stack(prews, 1)
