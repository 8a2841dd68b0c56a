"""TEST INFRASTRUCTURE ONLY (oracle): float64 restatement of the dense stacks of the reference's backup experiments.

Follows /root/reference src/distilation/backup/student_rollout_mlp_vf.py:251-276 (vf scope: tf.layers.dense stack, v_loss = reduce_sum(square(vpred - vtarg)))
and :608-616 (add_vtarg); backup/student_rollout.py:161-164,328 (reward head, squared-error term).  tf.layers.dense(x, n) = x @ W[in, n] + b[n].
Parity unpinned by reference artefacts (the backups store no vectors): pinned here by central finite differences (tests/test_dense_oracle.py).
Only tests/ may import this module.
"""
import numpy as np

from . import nn_np as NN


def unpack(dims, P):
    out, o = [], 0
    for i in range(len(dims) - 1):
        W = P[o:o + dims[i] * dims[i + 1]].reshape(dims[i], dims[i + 1]); o += dims[i] * dims[i + 1]
        b = P[o:o + dims[i + 1]]; o += dims[i + 1]
        out.append((W, b))
    assert o == P.size
    return out


def fwd(dims, acts, P, x):
    a = [np.asarray(x, np.float64)]
    for (W, b), act in zip(unpack(dims, np.asarray(P, np.float64)), acts):
        z = a[-1] @ W + b
        a.append(np.tanh(z) if act else z)
    return a


def loss_and_dout(s, t, kind):
    if kind == "mse":
        e = s - t
        return float(np.square(e).sum()), 2.0 * e
    return NN.kl_loss(s, t) if kind == "kl_st" else NN.kl_loss_rev(s, t)


def loss_grad(dims, acts, P, x, t, kind="mse"):
    """-> (outputs, loss, flat gradient)."""
    P = np.asarray(P, np.float64)
    a = fwd(dims, acts, P, x)
    loss, d = loss_and_dout(a[-1], np.asarray(t, np.float64).reshape(a[-1].shape), kind)
    layers = unpack(dims, P)
    g = []
    for l in range(len(layers) - 1, -1, -1):
        if acts[l]:
            d = d * (1.0 - a[l + 1] ** 2)
        g = [(a[l].T @ d).ravel(), d.sum(0)] + g
        d = d @ layers[l][0].T
    return a[-1], loss, np.concatenate(g)


def add_vtarg(reward_list, gamma, episode_steps=None):
    """Literal restatement of add_vtarg (:608-616) for ONE episode's reward list."""
    T = episode_steps or len(reward_list)
    target = np.empty(T, dtype=np.float64)
    target[-1] = pow(gamma, T) * reward_list[-1]
    for t in range(-2, -(T + 1), -1):
        target[t] = pow(gamma, t + T) * reward_list[t] + target[t + 1]
    return target
