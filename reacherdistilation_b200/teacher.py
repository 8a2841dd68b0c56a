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
    assert p.size == policy_param_count(nout)
    return p


def policy_param_count(nout):
    """rb_policy_param_count of the C ABI, restated on the host (ob_mean 11, ob_std 11, W1 b1 W2 b2 W3 b3, logstd 2) so that building a
    parameter vector needs no CUDA library (tests/test_abi.py checks the two agree)."""
    return 22 + 11 * 64 + 64 + 64 * 64 + 64 + 64 * nout + nout + 2


def policy_params_from_named(variables, nout=2, scope="pi"):
    """Flat C-ABI parameter vector from the variables of a baselines MlpPolicy checkpoint -- what `teacher.py:17-20` restores from
    `teacher.ckpt` (dump them once with `tf.train.load_checkpoint(...).get_tensor(name)` into a dict / npz and pass it here).
    Names as in the reference's graph (tfevents `pi/*`): `<scope>/obfilter/{runningsum, runningsumsq, count}`,
    `<scope>/pol/{fc1, fc2, final}/{w, b}` (`kernel` / `bias` accepted too), `<scope>/pol/logstd`; a trailing `:0` is ignored.
    obfilter: mean = runningsum / count, std = sqrt(max(runningsumsq / count - mean^2, 0.01))."""
    v = {k[:-2] if k.endswith(":0") else k: np.asarray(a, dtype=np.float64) for k, a in dict(variables).items()}

    def get(*names):
        for nm in names:
            if scope + "/" + nm in v:
                return v[scope + "/" + nm]
        raise KeyError("checkpoint has none of " + ", ".join(scope + "/" + nm for nm in names))

    count = float(np.ravel(get("obfilter/count"))[0])
    mean = np.ravel(get("obfilter/runningsum")) / count
    std = np.sqrt(np.maximum(np.ravel(get("obfilter/runningsumsq")) / count - mean * mean, 1e-2))
    parts = [mean, std]
    for layer, shape in (("fc1", (11, 64)), ("fc2", (64, 64)), ("final", (64, nout))):
        w, b = get("pol/%s/w" % layer, "pol/%s/kernel" % layer), get("pol/%s/b" % layer, "pol/%s/bias" % layer)
        if w.shape != shape or b.size != shape[1]:
            raise ValueError("%s/pol/%s: expected w %s, b [%d], got %s, %s" % (scope, layer, shape, shape[1], w.shape, b.shape))
        parts += [w.ravel(), np.ravel(b)]
    logstd = np.ravel(get("pol/logstd"))
    if logstd.size != 2 or mean.size != 11:
        raise ValueError("expected logstd [1,2] and an 11-d obfilter")
    parts.append(logstd)
    p = np.concatenate(parts).astype(np.float32)
    assert p.size == policy_param_count(nout)
    return p


def load_teacher_params(teacher_params=None, teacher_ckpt=None, nout=2):
    """The teacher weights a run distils (`teacher.py:17-20` restores teacher.ckpt; that file is not part of the reference repository).
    teacher_params: flat fp32 vector in the C-ABI layout.  teacher_ckpt: path of an .npz holding either `params` (flat) or the named
    variables of the baselines checkpoint (policy_params_from_named).  Returns (params, description); (None, None) when neither is given."""
    if teacher_params is not None:
        p = np.ascontiguousarray(np.asarray(teacher_params, dtype=np.float32).ravel())
        if p.size != policy_param_count(nout):
            raise ValueError("teacher_params has %d entries, the %d-output policy has %d" % (p.size, nout, policy_param_count(nout)))
        return p, "supplied parameter vector"
    if teacher_ckpt is not None:
        with np.load(teacher_ckpt) as z:
            if "params" in z.files:
                return load_teacher_params(z["params"], None, nout)[0], "flat parameters of %s" % teacher_ckpt
            return policy_params_from_named({k: z[k] for k in z.files}, nout=nout), "named variables of %s" % teacher_ckpt
    return None, None


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
            if restore:          # the reference ALWAYS restores teacher.ckpt (teacher.py:17-20): never substitute random weights for it silently
                raise _lib.ReacherB200Error("TeacherAgent(restore=True) needs weights: teacher.ckpt is not part of the reference repository -- pass "
                                            "params= (teacher.load_teacher_params / policy_params_from_named)")
            params = init_policy_params(seed=seed, nout=nout)     # untrained baselines-initialised teacher, seeded (synthetic workloads, tests)
        self.params_host = np.ascontiguousarray(params, dtype=np.float32)
        self.params = torch.from_numpy(self.params_host).to(self.device)
        self.pi = _Pi(self)
        # host-surface env at small batch (the reference's own shape: one env): the policy is evaluated by the env's resident server warp --
        # no launch, no copies per call, bit-identical to rb_policy_fwd(RB_MODE_FP32) (csrc/serve.cu)
        self._served_env = env if (env is not None and hasattr(env, "serve_policy") and env.serve_policy(self, self.params_host, nout)) else None
        self._pd_host = np.empty((env.num_envs, 4), np.float32) if self._served_env is not None else None

    def pdflat(self, ob, out=None):
        """sess.run(pi.pd.flat): ob [B,11] CUDA fp32 -> [B,4] (mean, logstd).  Host observations of a served env (numpy in -> numpy out): the
        pdflat the env server computed with the step that produced `ob`, else one server round trip."""
        env = self._served_env
        if env is not None and not torch.is_tensor(ob) and env._policy is self:
            if ob is env._last_ob:
                return env._last_pd
            o = np.ascontiguousarray(np.asarray(ob, dtype=np.float32).reshape(-1, 11))
            if o.shape[0] == env.num_envs:
                check(lib().rb_env_serve_policy_fwd(env._h, o.ctypes.data, self._pd_host.ctypes.data))
                return self._pd_host.copy()
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
