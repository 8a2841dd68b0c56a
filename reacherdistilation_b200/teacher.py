"""TeacherAgent: frozen baselines MlpPolicy (11 -> 64 tanh -> 64 tanh -> 2, obfilter clip +-5, state-independent logstd).

Mirrors /root/reference src/distilation/teacher.py:12-20.  The reference restores `teacher.ckpt`, which is not in the
repository, so weights are either supplied (`params=`) or drawn with the baselines initialisers (normc(1.0) hidden,
normc(0.01) final, zero biases) and the logstd recorded in the reference fixture.
Flat parameter layout: include/reacher_b200.h (rb_policy_param_count).
"""
import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr
from .config import TEACHER_LOGSTD


def normc(rng, shape, std):
    w = rng.standard_normal(shape)
    return (w * std / np.sqrt(np.square(w).sum(axis=0, keepdims=True))).astype(np.float32)


def init_policy_params(seed=0, nout=2, logstd=TEACHER_LOGSTD, ob_mean=None, ob_std=None, final_std=0.01):
    """baselines MlpPolicy initialisation, flat fp32 vector in the C-ABI layout."""
    rng = np.random.default_rng(seed)
    mu = np.zeros(11, np.float32) if ob_mean is None else np.asarray(ob_mean, np.float32)
    sd = np.ones(11, np.float32) if ob_std is None else np.asarray(ob_std, np.float32)
    parts = [mu, sd, normc(rng, (11, 64), 1.0).ravel(), np.zeros(64, np.float32), normc(rng, (64, 64), 1.0).ravel(), np.zeros(64, np.float32),
             normc(rng, (64, nout), final_std).ravel(), np.zeros(nout, np.float32), np.asarray(logstd, np.float32)]
    p = np.concatenate(parts).astype(np.float32)
    assert p.size == lib().rb_policy_param_count(nout)
    return p


class _Pd:
    def __init__(self, agent):
        self._a = agent


class _Pi:
    """Just enough of MlpPolicy for the loops: pi.act(stochastic=False, ob) and the (mean, flat) query."""

    def __init__(self, agent):
        self._a = agent

    def act(self, stochastic, ob):
        assert not stochastic, "the reference only ever acts deterministically (pd.mean)"
        flat = self._a.pdflat(ob)
        return flat[..., :2], None


class TeacherAgent:
    def __init__(self, env=None, sess=None, restore=False, batch=1, params=None, seed=0, device=None, mode=_lib.MODE_FP32, nout=2):
        dev = device if device is not None else (env.device if env is not None and hasattr(env, "device") else torch.device("cuda", 0))
        self.device = torch.device(dev)
        self.nout, self.mode = nout, mode
        if params is None:
            if restore:
                print("teacher.ckpt is not part of the reference repository; using seeded baselines-initialised weights")
            params = init_policy_params(seed=seed, nout=nout)
        self.params_host = np.ascontiguousarray(params, dtype=np.float32)
        self.params = torch.from_numpy(self.params_host).to(self.device)
        self.pi = _Pi(self)

    def pdflat(self, ob, out=None):
        """sess.run(pi.pd.flat): ob [B,11] CUDA fp32 -> [B,4] (mean, logstd)."""
        if not torch.is_tensor(ob):
            ob = torch.as_tensor(np.asarray(ob, dtype=np.float32).reshape(-1, 11)).to(self.device)
        ob = ob.reshape(-1, 11).contiguous()
        if out is None:
            out = torch.empty((ob.shape[0], 4), dtype=torch.float32, device=self.device)
        check(lib().rb_policy_fwd(ptr(self.params), self.nout, ptr(ob), ob.shape[0], ptr(out), self.mode, stream_ptr()))
        return out

    def mean_and_flat(self, ob):
        """sess.run((pi.pd.mean, pi.pd.flat)) -- mlp_train.py:123-125."""
        flat = self.pdflat(ob)
        return flat[:, :2], flat
