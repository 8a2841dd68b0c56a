"""ORACLE (test infrastructure, not product code) -- numpy restatement of teacher / students / KL / Adam / DAgger loop.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

Follows (reference = /root/reference):
  policy_fwd        teacher.py:12-16 (baselines MlpPolicy 'pi', hid 64x2 tanh; obfilter + clip +-5; op graph pinned by
                    the tfevents graph_defs under src/~/reacher/data/viz/1/) ; queried mlp_train.py:123-125,165-167.
                    Student variant a5' = backup/student_rollout.py:79-87 (gaussian_fixed_var=False -> 4 outputs).
  mlp_fwd/mlp_bwd   student_nn.py:51-57 (16 -> 24 tanh -> 128 tanh -> 128 linear -> 32 tanh -> 4), built mlp_train.py:50-55
  student_input     mlp_train.py:50-52  concat(dropout(ob, kp), prev_pdflat, prev_rew); tf.nn.dropout = x/kp*floor(kp+u)
  kl_loss           loss.py:3-13  (sum over everything, not mean)
  adam_tf           mlp_train.py:75-80 (tf.train.AdamOptimizer, epsilon OUTSIDE the bias correction: "epsilon hat")
                    backup/student_rollout.py:658,709 (MpiAdam(eps=1e-3).update(g, 1e-3): same formula on the averaged grad)
  dagger_iteration  mlp_train.py:143-204 / backup/student_rollout.py:682-709 restated for lock-step batches
                    (one env step of all envs + one optimiser step on the N fresh samples).
Parity status: KL formula pinned by fixture-derived KAT-8 (tests/test_oracle_fixture.py); teacher OP GRAPH pinned by
tfevents; teacher WEIGHTS and any loss curve are NOT in the reference ("parity unpinned": teacher.ckpt absent, no scalar
summaries) -> device parity for these is against this restatement with identical weights / batches.
"""
import numpy as np

from .philox_np import philox4x32_10
from .reacher_np import STREAM_DROPOUT

OB, PD, HID = 11, 4, 64


# ------------------------------------------------------------------ policy (MlpPolicy 11-64-64-nout) ----------
def policy_param_count(nout):
    return 22 + 11 * HID + HID + HID * HID + HID + HID * nout + nout + 2


def policy_unpack(p, nout):
    p = np.asarray(p)
    o = 0
    def take(n, shape=None):
        nonlocal o
        a = p[o:o + n]
        o += n
        return a.reshape(shape) if shape else a
    d = dict(mu=take(11), sd=take(11), W1=take(11 * HID, (11, HID)), b1=take(HID), W2=take(HID * HID, (HID, HID)), b2=take(HID),
             W3=take(HID * nout, (HID, nout)), b3=take(nout), logstd=take(2))
    assert o == policy_param_count(nout)
    return d


def policy_fwd(obs, params, nout=2, dtype=np.float64, return_hidden=False):
    """obs [B,11] -> pdflat [B,4].  nout=2: (mean, logstd broadcast); nout=4: raw 4 outputs."""
    P = {k: v.astype(dtype) for k, v in policy_unpack(np.asarray(params, dtype=np.float32), nout).items()}
    x = np.asarray(obs, dtype=dtype)
    z = np.clip((x - P["mu"]) / P["sd"], -5.0, 5.0)
    h1 = np.tanh(z @ P["W1"] + P["b1"])
    h2 = np.tanh(h1 @ P["W2"] + P["b2"])
    out = h2 @ P["W3"] + P["b3"]
    if nout == 2:
        out = np.concatenate([out, np.broadcast_to(P["logstd"], out.shape)], -1)
    return (out, z, h1, h2) if return_hidden else out


def policy_bwd(obs, params, dpd, dtype=np.float64):
    """Gradient of sum(dpd * pdflat) wrt the flat nout=4 policy parameters (obfilter / logstd slots get zero)."""
    nout = 4
    P = {k: v.astype(dtype) for k, v in policy_unpack(np.asarray(params, dtype=np.float32), nout).items()}
    out, z, h1, h2 = policy_fwd(obs, params, nout, dtype, return_hidden=True)
    d3 = np.asarray(dpd, dtype=dtype)
    gW3, gb3 = h2.T @ d3, d3.sum(0)
    d2 = (d3 @ P["W3"].T) * (1 - h2 * h2)
    gW2, gb2 = h1.T @ d2, d2.sum(0)
    d1 = (d2 @ P["W2"].T) * (1 - h1 * h1)
    gW1, gb1 = z.T @ d1, d1.sum(0)
    return np.concatenate([np.zeros(22, dtype), gW1.ravel(), gb1, gW2.ravel(), gb2, gW3.ravel(), gb3, np.zeros(2, dtype)])


# ------------------------------------------------------------------ generic student MLP (student_nn.py:51-57) --
MLP_DIMS = (16, 24, 128, 128, 32, 4)
MLP_TANH = (True, True, False, True, False)


def mlp_param_count(dims=MLP_DIMS):
    return sum(dims[i] * dims[i + 1] + dims[i + 1] for i in range(len(dims) - 1))


def mlp_unpack(p, dims=MLP_DIMS):
    out, o = [], 0
    for i in range(len(dims) - 1):
        W = p[o:o + dims[i] * dims[i + 1]].reshape(dims[i], dims[i + 1]); o += W.size
        b = p[o:o + dims[i + 1]]; o += b.size
        out.append((W, b))
    assert o == mlp_param_count(dims)
    return out


def mlp_fwd(x, params, dims=MLP_DIMS, acts=MLP_TANH, dtype=np.float64):
    layers = mlp_unpack(np.asarray(params, dtype=np.float32).astype(dtype), dims)
    hs = [np.asarray(x, dtype=dtype)]
    for (W, b), a in zip(layers, acts):
        y = hs[-1] @ W + b
        hs.append(np.tanh(y) if a else y)
    return hs[-1], hs


def mlp_bwd(hs, params, dout, dims=MLP_DIMS, acts=MLP_TANH, dtype=np.float64):
    layers = mlp_unpack(np.asarray(params, dtype=np.float32).astype(dtype), dims)
    g = [None] * len(layers)
    d = np.asarray(dout, dtype=dtype)
    for i in reversed(range(len(layers))):
        W, _ = layers[i]
        if acts[i]:
            d = d * (1 - hs[i + 1] * hs[i + 1])
        g[i] = np.concatenate([(hs[i].T @ d).ravel(), d.sum(0)])
        d = d @ W.T
    return np.concatenate(g)


def dropout_keep(seed, sample_ids, iteration, keep_prob, width=OB):
    """Philox dropout keep-mask, float32 rule floor(kp + u) with u=(x>>8)*2^-24; counter=(sample, iteration, word/4, 2)."""
    sample_ids = np.asarray(sample_ids, dtype=np.uint32)
    it = np.full_like(sample_ids, np.uint32(iteration))
    cols = []
    for blk in range((width + 3) // 4):
        r = philox4x32_10(seed, sample_ids, it, np.full_like(sample_ids, blk), np.full_like(sample_ids, STREAM_DROPOUT))
        cols.extend(r)
    u = np.stack(cols[:width], -1)
    uf = (u >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
    return np.floor(np.float32(keep_prob) + uf).astype(np.float32)


def student_input(obs, prev_pdflat, prev_rew, keep_prob, seed, sample_ids, iteration, dtype=np.float64):
    """mlp_train.py:50-52.  obs is float32-valued; dropout scaling done in float32 like TF."""
    ob32 = np.asarray(obs, dtype=np.float32)
    if keep_prob < 1.0:
        keep = dropout_keep(seed, sample_ids, iteration, keep_prob)
        ob32 = (ob32 / np.float32(keep_prob)) * keep
    return np.concatenate([ob32.astype(dtype), np.asarray(prev_pdflat, dtype), np.asarray(prev_rew, dtype).reshape(-1, 1)], -1)


# ------------------------------------------------------------------ loss (loss.py:3-13) ------------------------
def kl_loss(s, t, dtype=np.float64):
    """KL(student || teacher) of diagonal Gaussians, SUM over all samples and dims.  Returns (loss, dL/ds)."""
    s, t = np.asarray(s, dtype=dtype), np.asarray(t, dtype=dtype)
    ms, ls, mt, lt = s[..., :2], s[..., 2:], t[..., :2], t[..., 2:]
    vs, vt = np.exp(2 * ls), np.exp(2 * lt)
    per = lt - ls + (vs + (ms - mt) ** 2) / (2 * vt) - 0.5
    grad = np.concatenate([(ms - mt) / vt, -1.0 + vs / vt], -1)
    return per.sum(), grad


def kl_loss_rev(s, t, dtype=np.float64):
    """KL(teacher || student) (backup/student_rollout.py:639-640 `klts`).  Returns (loss, dL/ds)."""
    s, t = np.asarray(s, dtype=dtype), np.asarray(t, dtype=dtype)
    ms, ls, mt, lt = s[..., :2], s[..., 2:], t[..., :2], t[..., 2:]
    vs, vt = np.exp(2 * ls), np.exp(2 * lt)
    per = ls - lt + (vt + (ms - mt) ** 2) / (2 * vs) - 0.5
    grad = np.concatenate([(ms - mt) / vs, 1.0 - (vt + (ms - mt) ** 2) / vs], -1)
    return per.sum(), grad


def mse_loss(s, t, actions_only=False, dtype=np.float64):
    """Squared-error distillation (north_star (2) "MSE/KL on actions"; backup/student_rollout.py:328 uses the same sum-of-squares form for
    its reward head): SUM of (s - t)^2 over the pdflat row, or over its mean half only (the actions).  Returns (loss, dL/ds)."""
    s, t = np.asarray(s, dtype=dtype), np.asarray(t, dtype=dtype)
    e = s - t
    if actions_only:
        e = np.concatenate([e[..., :2], np.zeros_like(e[..., 2:])], -1)
    return (e * e).sum(), 2.0 * e


def pd_loss(s, t, kind, dtype=np.float64):
    """Loss by the C-ABI constant: 0 KL(s||t), 1 KL(t||s), 2 squared error on pdflat, 3 squared error on the actions."""
    return (kl_loss, kl_loss_rev, mse_loss, lambda a, b, dtype=np.float64: mse_loss(a, b, True, dtype))[kind](s, t, dtype=dtype)


# ------------------------------------------------------------------ optimiser ---------------------------------
class AdamTF:
    """tf.train.AdamOptimizer / baselines MpiAdam update rule (epsilon added to the UNcorrected sqrt(v))."""

    def __init__(self, n, lr=1e-4, beta1=0.9, beta2=0.999, eps=1e-8, dtype=np.float64):
        self.m, self.v, self.t = np.zeros(n, dtype), np.zeros(n, dtype), 0
        self.lr, self.b1, self.b2, self.eps, self.dtype = lr, beta1, beta2, eps, dtype

    def update(self, theta, g):
        self.t += 1
        g = np.asarray(g, dtype=self.dtype)
        lr_t = self.lr * np.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        self.m = self.b1 * self.m + (1 - self.b1) * g
        self.v = self.b2 * self.v + (1 - self.b2) * g * g
        return theta - lr_t * self.m / (np.sqrt(self.v) + self.eps)


# ------------------------------------------------------------------ initialisers (baselines normc / tf glorot) --
def normc_init(rng, shape, std):
    w = rng.standard_normal(shape)
    return (w * std / np.sqrt(np.square(w).sum(0, keepdims=True))).astype(np.float32)
