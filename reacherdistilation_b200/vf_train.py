"""Auxiliary objectives of the reference's backup experiments on the device (rb_dense_* / rb_vf_targets, csrc/dense.cu).

* `DenseNet`            generic dense stack trained with a summed squared error or the KL losses (tf.layers.dense stacks of the backups)
* `ValueRegressor`      the "vf" scope of src/distilation/backup/student_rollout_mlp_vf.py:251-276: [prev_ob | next_ac] -> 64 (linear) ->
                        10 x tanh(100) -> 1, v_loss = sum (vpred - vtarg)^2, Adam(lr 1e-2) (:290-295)
* `RewardHead`          src/distilation/backup/student_rollout.py:161-164,328: tanh(dense 64) -> dense 1 on trunk features, summed squared error
* `add_vtarg`           backup/student_rollout_mlp_vf.py:608-616, batched over episodes
"""
import ctypes as C

import numpy as np
import torch

from ._lib import LOSS_KL_ST, LOSS_MSE, check, lib, ptr, stream_ptr

GAMMA = 0.99            # backup/student_rollout_mlp_vf.py:41
EPISODE_STEPS = 50


def normc_initializer(rng, shape, std=1.0):
    """baselines U.normc_initializer: N(0,1) columns scaled to norm `std`."""
    w = rng.standard_normal(shape)
    return (w * std / np.sqrt(np.square(w).sum(0, keepdims=True))).astype(np.float32)


def init_dense_params(dims, seed=0, init="glorot"):
    """Flat [W_l (in x out, row-major), b_l] per layer.  glorot: tf.layers.dense default; normc: the vf scope's kernel_initializer."""
    rng = np.random.default_rng(seed)
    parts = []
    for i in range(len(dims) - 1):
        if init == "normc":
            w = normc_initializer(rng, (dims[i], dims[i + 1]), 1.0)
        else:
            lim = np.sqrt(6.0 / (dims[i] + dims[i + 1]))
            w = rng.uniform(-lim, lim, size=(dims[i], dims[i + 1])).astype(np.float32)
        parts += [w.ravel(), np.zeros(dims[i + 1], np.float32)]
    return np.concatenate(parts)


class DenseNet:
    """Owns the flat parameters, Adam moments, [grad | loss] vector and workspace of one dense stack."""

    def __init__(self, dims, acts=None, seed=0, device=0, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, params=None, init="glorot"):
        self.dims = [int(d) for d in dims]
        self.L = len(self.dims) - 1
        self.acts = [int(a) for a in acts] if acts is not None else [1] * (self.L - 1) + [0]
        assert len(self.acts) == self.L
        self._dims_c = (C.c_int * (self.L + 1))(*self.dims)
        self._acts_c = (C.c_int * self.L)(*self.acts)
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self.P = int(lib().rb_dense_param_count(self.L, self._dims_c))
        if self.P <= 0:
            raise ValueError("bad dense stack %r" % (self.dims,))
        if params is None:
            params = init_dense_params(self.dims, seed, init)
        assert params.size == self.P
        self.lr, self.beta1, self.beta2, self.eps = lr, beta1, beta2, eps
        with torch.cuda.device(self.device):
            self.params = torch.from_numpy(np.ascontiguousarray(params, np.float32)).to(self.device)
            self.m, self.v = torch.zeros_like(self.params), torch.zeros_like(self.params)
            self.gradloss = torch.zeros(self.P + 1, dtype=torch.float32, device=self.device)
        self.t, self._ws, self._ws_batch = 0, None, 0

    def _workspace(self, B):
        if B > self._ws_batch:
            nbytes = int(lib().rb_dense_workspace_bytes(self.L, self._dims_c, B))
            with torch.cuda.device(self.device):
                self._ws = torch.empty(nbytes // 4, dtype=torch.float32, device=self.device)
            self._ws_batch = B
        return self._ws

    def forward(self, x, out=None):
        x = x.contiguous()
        B = x.shape[0]
        assert x.shape[1] == self.dims[0]
        if out is None:
            out = torch.empty((B, self.dims[-1]), dtype=torch.float32, device=self.device)
        check(lib().rb_dense_fwd(ptr(self.params), self.L, self._dims_c, self._acts_c, ptr(x), B, ptr(out), ptr(self._workspace(B)), stream_ptr()))
        return out

    def loss_grad(self, x, target, loss_kind=LOSS_MSE, out=None):
        """Fills self.gradloss = [flat gradient | loss]; returns the outputs [B, dims[-1]]."""
        x, target = x.contiguous(), target.contiguous()
        B = x.shape[0]
        assert x.shape[1] == self.dims[0] and target.numel() == B * self.dims[-1]
        if out is None:
            out = torch.empty((B, self.dims[-1]), dtype=torch.float32, device=self.device)
        check(lib().rb_dense_loss_grad(ptr(self.params), self.L, self._dims_c, self._acts_c, ptr(x), ptr(target), B, loss_kind, ptr(out),
                                       ptr(self.gradloss), ptr(self._workspace(B)), stream_ptr()))
        return out

    def last_loss(self):
        return self.gradloss[self.P]

    def adam_step(self, grad_scale=1.0):
        self.t += 1
        check(lib().rb_adam_step(ptr(self.params), ptr(self.m), ptr(self.v), ptr(self.gradloss), self.P, self.t, self.lr, self.beta1, self.beta2,
                                 self.eps, grad_scale, stream_ptr()))

    def step(self, x, target, loss_kind=LOSS_MSE):
        out = self.loss_grad(x, target, loss_kind)
        self.adam_step()
        return out


class ValueRegressor(DenseNet):
    """vpred([prev_ob | next_ac]) trained on add_vtarg targets (backup/student_rollout_mlp_vf.py:251-276, 290-295)."""
    DIMS = (13, 64) + (100,) * 10 + (1,)
    ACTS = (0,) + (1,) * 10 + (0,)

    def __init__(self, seed=0, device=0, lr=1e-2, **kw):
        super().__init__(self.DIMS, self.ACTS, seed=seed, device=device, lr=lr, init="normc", **kw)

    def fit_step(self, prev_ob, next_ac, vtarg):
        """One minimize_vf_l2 step on [B, T-1, .] tensors (the reference's placeholders, :253-259)."""
        x = torch.cat([prev_ob.reshape(-1, 11), next_ac.reshape(-1, 2)], -1)
        return self.step(x, vtarg.reshape(-1, 1), LOSS_MSE)


class RewardHead(DenseNet):
    """reward = dense1(tanh(dense64(features))) with a summed squared error (backup/student_rollout.py:161-164,328)."""

    def __init__(self, in_dim=128, seed=0, device=0, lr=1e-3, **kw):
        super().__init__((in_dim, 64, 1), (1, 0), seed=seed, device=device, lr=lr, **kw)


def add_vtarg(reward, gamma=GAMMA, out=None):
    """reward [E, T] (device) -> vtarg [E, T]: target[T-1] = gamma^T r[T-1]; target[i] = gamma^i r[i] + target[i+1]."""
    reward = reward.contiguous()
    E, T = reward.shape
    if out is None:
        out = torch.empty_like(reward)
    check(lib().rb_vf_targets(ptr(reward), E, T, float(gamma), ptr(out), stream_ptr()))
    return out
