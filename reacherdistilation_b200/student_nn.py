"""Student networks (/root/reference src/distilation/student_nn.py:51-57 and backup/student_rollout.py:79-87).

StudentNet owns the flat fp32 parameter vector, Adam moments and the workspace on one GPU and drives the CUDA
forward / loss+grad / Adam entry points.  `student_mlp_graph` keeps the reference's function name for the forward pass.
"""
import numpy as np
import torch

from . import _lib
from ._lib import LOSS_KL_ST, MODE_FP32, STUDENT_MLP, STUDENT_POLICY64, check, lib, ptr, stream_ptr
from .teacher import init_policy_params

MLP_DIMS = (16, 24, 128, 128, 32, 4)


def init_mlp_params(seed=0, dims=MLP_DIMS):
    """tf.layers.dense defaults: glorot-uniform kernels, zero biases; flat layout W[in][out], b[out] per layer."""
    rng = np.random.default_rng(seed)
    parts = []
    for i in range(len(dims) - 1):
        lim = np.sqrt(6.0 / (dims[i] + dims[i + 1]))
        parts.append(rng.uniform(-lim, lim, size=(dims[i], dims[i + 1])).astype(np.float32).ravel())
        parts.append(np.zeros(dims[i + 1], np.float32))
    return np.concatenate(parts)


class StudentNet:
    def __init__(self, kind=STUDENT_MLP, seed=0, device=0, lr=None, beta1=0.9, beta2=0.999, eps=None, mode=MODE_FP32, params=None):
        self.kind, self.mode = kind, mode
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self.P = lib().rb_student_param_count(kind)
        self.in_dim = lib().rb_student_input_dim(kind)
        if params is None:
            params = init_mlp_params(seed) if kind == STUDENT_MLP else init_policy_params(seed, nout=4, logstd=(0.0, 0.0))
        assert params.size == self.P
        # reference optimiser settings: tf Adam lr 1e-4 eps 1e-8 (mlp_train.py:75-78); MpiAdam eps 1e-3 step 1e-3 (backup :658,709)
        self.lr = lr if lr is not None else (1e-4 if kind == STUDENT_MLP else 1e-3)
        self.eps = eps if eps is not None else (1e-8 if kind == STUDENT_MLP else 1e-3)
        self.beta1, self.beta2 = beta1, beta2
        with torch.cuda.device(self.device):
            self.params = torch.from_numpy(np.ascontiguousarray(params, np.float32)).to(self.device)
            self.m = torch.zeros_like(self.params)
            self.v = torch.zeros_like(self.params)
            self.gradloss = torch.zeros(self.P + 1, dtype=torch.float32, device=self.device)
            ws = lib().rb_student_workspace_bytes(kind, 0, self.device.index)
            self.workspace = torch.empty(ws // 4, dtype=torch.float32, device=self.device)
        self.t = 0

    def forward(self, x, out=None):
        x = x.reshape(-1, self.in_dim).contiguous()
        if out is None:
            out = torch.empty((x.shape[0], 4), dtype=torch.float32, device=self.device)
        check(lib().rb_student_fwd_ws(self.kind, ptr(self.params), ptr(x), x.shape[0], ptr(out), ptr(self.workspace), self.mode, stream_ptr()))
        return out

    def loss_grad(self, x, t_pdflat, loss_kind=LOSS_KL_ST, s_out=None):
        """Fills self.gradloss = [flat grad (P), summed loss]; returns the student pdflat [B,4]."""
        x = x.reshape(-1, self.in_dim).contiguous()
        t_pdflat = t_pdflat.reshape(-1, 4).contiguous()
        B = x.shape[0]
        if s_out is None:
            s_out = torch.empty((B, 4), dtype=torch.float32, device=self.device)
        check(lib().rb_student_loss_grad(self.kind, ptr(self.params), ptr(x), ptr(t_pdflat), B, loss_kind, ptr(s_out), ptr(self.gradloss),
                                         ptr(self.workspace), self.mode, stream_ptr()))
        return s_out

    def step(self, x, t_pdflat, loss_kind=LOSS_KL_ST, s_out=None, grad_scale=1.0):
        """loss_grad + adam_step in one C-ABI call (single rank); RB_MODE_TC: one cooperative kernel launch."""
        x = x.reshape(-1, self.in_dim).contiguous()
        t_pdflat = t_pdflat.reshape(-1, 4).contiguous()
        B = x.shape[0]
        if s_out is None:
            s_out = torch.empty((B, 4), dtype=torch.float32, device=self.device)
        self.t += 1
        check(lib().rb_student_step(self.kind, ptr(self.params), ptr(self.m), ptr(self.v), ptr(x), ptr(t_pdflat), B, loss_kind, ptr(s_out),
                                    ptr(self.gradloss), ptr(self.workspace), self.t, self.lr, self.beta1, self.beta2, self.eps, grad_scale,
                                    self.mode, stream_ptr()))
        return s_out

    # ---- data parallel: one-shot all-reduce over NVLink peer memory fused into the step kernel -------------------------
    def enable_peer_exchange(self, group=None):
        """Allocate this rank's symmetric buffer [reserved(64 u32) | receive area 0 | receive area 1] (an area = {value, epoch} pairs
        [world][slot], see rb_student_step_dp) and exchange the peer mappings (torch.distributed._symmetric_memory: CUDA VMM handles over
        the process group's store).  RB_MODE_TC only."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        assert self.mode == _lib.MODE_TC, "the fused exchange lives in the tcgen05 student kernel"
        group = group if group is not None else dist.group.WORLD
        self._px_world, self._px_rank = dist.get_world_size(group), dist.get_rank(group)
        assert 2 <= self._px_world <= 8
        slot = (self.P + 1 + 63) // 64 * 64
        area = 2 * self._px_world * slot                # floats: 8-byte pairs, one row per sending rank
        with torch.cuda.device(self.device):
            self._px_buf = symm.empty(64 + 2 * area, dtype=torch.float32, device=self.device)
        self._px_buf.zero_()
        hdl = symm.rendezvous(self._px_buf, group.group_name)
        self._px_hdl = hdl
        base = [int(p) for p in hdl.buffer_ptrs]
        mk = lambda vals: torch.tensor(vals, dtype=torch.int64).numpy().astype(np.uint64)
        self._px_flags = mk(base)
        self._px_slots = [mk([b + 4 * (64 + k * area) for b in base]) for k in (0, 1)]
        self._px_epoch = 0
        torch.cuda.synchronize(self.device)
        hdl.barrier()                                   # every rank's receive areas are zero before anyone steps
        torch.cuda.synchronize(self.device)

    def step_dp(self, x, t_pdflat, loss_kind=LOSS_KL_ST, s_out=None, grad_scale=1.0):
        """Data-parallel rb_student_step: local [grad|loss] -> all-reduce over peer memory -> Adam, in one kernel per rank."""
        x = x.reshape(-1, self.in_dim).contiguous()
        t_pdflat = t_pdflat.reshape(-1, 4).contiguous()
        B = x.shape[0]
        if s_out is None:
            s_out = torch.empty((B, 4), dtype=torch.float32, device=self.device)
        self.t += 1
        self._px_epoch += 1
        slots = self._px_slots[self._px_epoch & 1]
        check(lib().rb_student_step_dp(self.kind, ptr(self.params), ptr(self.m), ptr(self.v), ptr(x), ptr(t_pdflat), B, loss_kind, ptr(s_out),
                                       ptr(self.gradloss), ptr(self.workspace), self.t, self.lr, self.beta1, self.beta2, self.eps, grad_scale,
                                       self._px_rank, self._px_world, slots.ctypes.data, self._px_flags.ctypes.data, self._px_epoch, stream_ptr()))
        return s_out

    def adam_step(self, grad_scale=1.0):
        self.t += 1
        check(lib().rb_adam_step(ptr(self.params), ptr(self.m), ptr(self.v), ptr(self.gradloss), self.P, self.t, self.lr, self.beta1, self.beta2,
                                 self.eps, grad_scale, stream_ptr()))

    def state_dict(self):
        return dict(kind=self.kind, params=self.params.cpu(), m=self.m.cpu(), v=self.v.cpu(), t=self.t)

    def load_state_dict(self, sd):
        assert sd["kind"] == self.kind
        self.params.copy_(sd["params"]); self.m.copy_(sd["m"]); self.v.copy_(sd["v"]); self.t = int(sd["t"])


def student_mlp_input(obs, prev_pdflat, prev_rew, keep_prob, seed, sample_id0=0, iteration=0, out=None):
    """concat(dropout(ob, kp), prev_pdflat, prev_rew) -- mlp_train.py:50-52."""
    B = obs.shape[0]
    if out is None:
        out = torch.empty((B, 16), dtype=torch.float32, device=obs.device)
    check(lib().rb_student_mlp_input(ptr(obs.contiguous()), ptr(prev_pdflat), ptr(prev_rew), B, float(keep_prob), int(seed), int(sample_id0),
                                     int(iteration), ptr(out), stream_ptr()))
    return out


def student_mlp_graph(training_input_batch, net):
    """Reference name (student_nn.py:51): forward of the 16-24-128-128-32-4 student on [.., 16] inputs -> [.., 4]."""
    shp = training_input_batch.shape[:-1]
    return net.forward(training_input_batch).reshape(*shp, 4)


class StudentLSTM:
    """student_lstm_graph (student_nn.py:21-49): owns the flat parameters, Adam moments and workspace of the LSTM student.
    loss_grad / forward take time-major windows [T=10, B, .]; the carried acting state is [2, B, 200] (c, m)."""

    def __init__(self, seed=0, device=0, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, params=None, max_batch=None):
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self.P, self.T, self.U = int(lib().rb_lstm_param_count()), int(lib().rb_lstm_steps()), int(lib().rb_lstm_units())
        if params is None:
            params = init_lstm_params(seed)
        assert params.size == self.P
        self.lr, self.beta1, self.beta2, self.eps = lr, beta1, beta2, eps          # lstm_train.py:74-78
        with torch.cuda.device(self.device):
            self.params = torch.from_numpy(np.ascontiguousarray(params, np.float32)).to(self.device)
            self.m, self.v = torch.zeros_like(self.params), torch.zeros_like(self.params)
            self.gradloss = torch.zeros(self.P + 1, dtype=torch.float32, device=self.device)
        self.t, self._ws, self._ws_batch = 0, None, 0
        if max_batch:
            self._workspace(max_batch)

    def _workspace(self, B):
        if B > self._ws_batch:
            nbytes = int(lib().rb_lstm_workspace_bytes(B))
            with torch.cuda.device(self.device):
                self._ws = torch.empty(nbytes // 4, dtype=torch.float32, device=self.device)
            self._ws_batch = B
        return self._ws

    def zero_state(self, B):
        return torch.zeros((2, B, self.U), dtype=torch.float32, device=self.device)

    def forward(self, ob, prev_pdflat, state=None):
        """Acting (lstm_train.py:171-182): -> (s_pdflat [T,B,4], final_state [2,B,200])."""
        T, B = ob.shape[0], ob.shape[1]
        assert T == self.T
        ob, prev_pdflat = ob.contiguous(), prev_pdflat.contiguous()
        s = torch.empty((T, B, 4), dtype=torch.float32, device=self.device)
        fin = torch.empty((2, B, self.U), dtype=torch.float32, device=self.device)
        check(lib().rb_lstm_fwd(ptr(self.params), ptr(ob), ptr(prev_pdflat), ptr(state.contiguous() if state is not None else None), B, ptr(s), ptr(fin),
                                ptr(self._workspace(B)), stream_ptr()))
        return s, fin

    def loss_grad(self, ob, prev_pdflat, t_pdflat, state=None, keep_prob=1.0, seed=0, sample_id0=0, iteration=0, loss_kind=LOSS_KL_ST,
                  final_state_out=None):
        """Training window batch (lstm_train.py:145-160): fills self.gradloss = [flat grad | loss]; returns s_pdflat [T,B,4].
        final_state_out [2,B,200]: receives the (c, m) state after the last unrolled step -- what backup/lstm_bbpt.py:125-137 fetches as
        `final_state_batch` next to the loss and feeds into the next batch."""
        T, B = ob.shape[0], ob.shape[1]
        assert T == self.T
        ob, prev_pdflat, t_pdflat = ob.contiguous(), prev_pdflat.contiguous(), t_pdflat.contiguous()
        if final_state_out is not None:
            assert final_state_out.shape == (2, B, self.U) and final_state_out.is_contiguous() and final_state_out.dtype == torch.float32
        s = torch.empty((T, B, 4), dtype=torch.float32, device=self.device)
        check(lib().rb_lstm_loss_grad(ptr(self.params), ptr(ob), ptr(prev_pdflat), ptr(t_pdflat), ptr(state.contiguous() if state is not None else None), B,
                                      float(keep_prob), int(seed), int(sample_id0), int(iteration), loss_kind, ptr(s), ptr(final_state_out),
                                      ptr(self.gradloss), ptr(self._workspace(B)), stream_ptr()))
        return s

    def adam_step(self, grad_scale=1.0):
        self.t += 1
        check(lib().rb_adam_step(ptr(self.params), ptr(self.m), ptr(self.v), ptr(self.gradloss), self.P, self.t, self.lr, self.beta1, self.beta2,
                                 self.eps, grad_scale, stream_ptr()))

    def step(self, ob, prev_pdflat, t_pdflat, state=None, keep_prob=1.0, seed=0, sample_id0=0, loss_kind=LOSS_KL_ST, s_out=None, use_graph=True):
        """loss_grad + Adam as one CUDA-graph launch (rb_lstm_step).  The graph is re-captured whenever a pointer or scalar changes, so pass
        the same (static) input tensors every step; the dropout iteration and the Adam step count live in a device-side clock."""
        import ctypes as C
        T, B = ob.shape[0], ob.shape[1]
        assert T == self.T and ob.is_contiguous() and prev_pdflat.is_contiguous() and t_pdflat.is_contiguous()
        if getattr(self, "_ctx", None) is None:
            h = C.c_void_p()
            check(lib().rb_lstm_ctx_create(C.byref(h), self.device.index))
            self._ctx, self._clock_t = h, None
        if self._clock_t != self.t:                       # (re)load the device clock after any non-graph update
            check(lib().rb_lstm_ctx_set_clock(self._ctx, self.t, self.t, stream_ptr()))
        if s_out is None:                                   # static output buffer (a fresh tensor every step would re-capture the graph)
            cache = self.__dict__.setdefault("_s_static", {})
            if B not in cache:
                cache[B] = torch.empty((T, B, 4), dtype=torch.float32, device=self.device)
            s_out = cache[B]
        check(lib().rb_lstm_step(self._ctx, ptr(self.params), ptr(self.m), ptr(self.v), ptr(ob), ptr(prev_pdflat), ptr(t_pdflat),
                                 ptr(state.contiguous() if state is not None else None), B, float(keep_prob), int(seed), int(sample_id0), loss_kind,
                                 ptr(s_out), ptr(self.gradloss), ptr(self._workspace(B)), self.lr, self.beta1, self.beta2, self.eps, 1.0,
                                 1 if use_graph else 0, stream_ptr()))
        self.t += 1
        self._clock_t = self.t
        return s_out

    def state_dict(self):
        return dict(kind="lstm", params=self.params.cpu(), m=self.m.cpu(), v=self.v.cpu(), t=self.t)

    def load_state_dict(self, sd):
        assert sd["kind"] == "lstm"
        self.params.copy_(sd["params"]); self.m.copy_(sd["m"]); self.v.copy_(sd["v"]); self.t = int(sd["t"])


LSTM_HEAD_DIMS = (200, 64, 128, 64, 32, 4)


def init_lstm_params(seed=0):
    """glorot-uniform kernels, zero biases (tf.layers.dense and LSTMCell defaults); layout: include/reacher_b200.h."""
    rng = np.random.default_rng(seed)
    parts = []
    def glorot(fi, fo):
        lim = np.sqrt(6.0 / (fi + fo))
        parts.append(rng.uniform(-lim, lim, fi * fo).astype(np.float32)); parts.append(np.zeros(fo, np.float32))
    glorot(4, 32)
    glorot(243, 800)
    for _ in range(10):
        for l in range(5):
            glorot(LSTM_HEAD_DIMS[l], LSTM_HEAD_DIMS[l + 1])
    return np.concatenate(parts)


def student_lstm_graph(ob_batch, keep_prob, prev_pdflat_batch, initial_state_batch, net, seed=0, iteration=0):
    """Reference name (student_nn.py:21): forward of the LSTM student; keep_prob < 1 uses the Philox dropout mask."""
    if keep_prob >= 1.0:
        return net.forward(ob_batch, prev_pdflat_batch, initial_state_batch)
    raise NotImplementedError("dropout forward is part of loss_grad (training); acting uses keep_prob = 1 (lstm_train.py:176)")


# ---------------------------------------------------------------------------------------------- two-headed LSTM student (backup experiment)
def lstm2_spec(units=1, steps=2, carry_state=False, trunk=128, action_hidden=64, reward_hidden=(64,)):
    """spec of rb_lstm2_* (include/reacher_b200.h).  Defaults = the checked-in source, backup/student_rollout.py:38-50,130-172 (NUM_UNITS 1,
    STEPS_UNROLLED 2; its loop never reassigns `state`, so no state is carried); `LSTM2_TFEVENTS_SPEC` is the graph of the recorded tfevents."""
    a = np.zeros(10, np.int32)
    a[:6] = (units, steps, 1 if carry_state else 0, trunk, action_hidden, len(reward_hidden))
    a[6:6 + len(reward_hidden)] = reward_hidden
    return a


LSTM2_TFEVENTS_SPEC = dict(units=1, steps=2, carry_state=True, reward_hidden=(64, 32, 64))


def init_lstm2_params(spec, seed=0):
    """glorot-uniform kernels, zero biases; layout: include/reacher_b200.h (cell, then per step trunk | reward layers | reward_out | action | pd)."""
    U, T, _, D, A, nR = (int(v) for v in spec[:6])
    rh = [int(v) for v in spec[6:6 + nR]]
    rng = np.random.default_rng(seed)
    parts = []
    def glorot(fi, fo):
        lim = np.sqrt(6.0 / (fi + fo))
        parts.append(rng.uniform(-lim, lim, fi * fo).astype(np.float32)); parts.append(np.zeros(fo, np.float32))
    glorot(13 + U, 4 * U)
    for _ in range(T):
        glorot(U, D)
        prev = D
        for r in rh:
            glorot(prev, r)
            prev = r
        glorot(prev, 1)
        glorot(D, A)
        glorot(A, 4)
    return np.concatenate(parts)


class StudentLSTM2:
    """lstm_graph + lstm_loss + the reward term (backup/student_rollout.py:130-200,328): flat parameters, Adam moments (lr 1e-3, :331-336) and
    workspace of the two-headed LSTM student.  Windows are time-major [T,B,.]; state [2,B,units] (c, m)."""

    def __init__(self, spec=None, seed=0, device=0, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, params=None):
        import ctypes as C
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self.spec = np.ascontiguousarray(lstm2_spec() if spec is None else spec, np.int32)
        assert self.spec.size == 10
        self._spec_p = self.spec.ctypes.data_as(C.POINTER(C.c_int))
        self.P = int(lib().rb_lstm2_param_count(self._spec_p))
        if self.P < 0:
            raise _lib.ReacherB200Error(lib().rb_last_error().decode())
        self.U, self.T, self.carry_state = int(self.spec[0]), int(self.spec[1]), bool(self.spec[2])
        if params is None:
            params = init_lstm2_params(self.spec, seed)
        assert params.size == self.P
        self.lr, self.beta1, self.beta2, self.eps = lr, beta1, beta2, eps
        with torch.cuda.device(self.device):
            self.params = torch.from_numpy(np.ascontiguousarray(params, np.float32)).to(self.device)
            self.m, self.v = torch.zeros_like(self.params), torch.zeros_like(self.params)
            self.gradloss = torch.zeros(self.P + 3, dtype=torch.float32, device=self.device)       # grad | total | KL part | reward part
        self.t, self._ws, self._ws_batch = 0, None, 0

    def _workspace(self, B):
        if B > self._ws_batch:
            nbytes = int(lib().rb_lstm2_workspace_bytes(self._spec_p, B))
            with torch.cuda.device(self.device):
                self._ws = torch.empty(nbytes // 4, dtype=torch.float32, device=self.device)
            self._ws_batch = B
        return self._ws

    def zero_state(self, B):
        return torch.zeros((2, B, self.U), dtype=torch.float32, device=self.device)

    def forward(self, ob, action, state=None):
        """Acting (:527-535): -> (s_pdflat [T,B,4], reward [T,B], final_state [2,B,units])."""
        T, B = ob.shape[0], ob.shape[1]
        assert T == self.T and action.shape == (T, B, 2)
        ob, action = ob.contiguous(), action.contiguous()
        s = torch.empty((T, B, 4), dtype=torch.float32, device=self.device)
        rew = torch.empty((T, B), dtype=torch.float32, device=self.device)
        fin = torch.empty((2, B, self.U), dtype=torch.float32, device=self.device)
        check(lib().rb_lstm2_fwd(self._spec_p, ptr(self.params), ptr(ob), ptr(action), ptr(state.contiguous() if state is not None else None), B, ptr(s),
                                 ptr(rew), ptr(fin), ptr(self._workspace(B)), stream_ptr()))
        return s, rew, fin

    def loss_grad(self, ob, action, t_pdflat, reward_target, state=None, keep_prob=1.0, seed=0, sample_id0=0, iteration=0, loss_kind=LOSS_KL_ST):
        """Training windows (:508-522): fills self.gradloss = [flat grad | total loss | KL | reward sse]; returns (s_pdflat, reward)."""
        T, B = ob.shape[0], ob.shape[1]
        assert T == self.T and action.shape == (T, B, 2) and t_pdflat.shape == (T, B, 4) and reward_target.shape == (T, B)
        ob, action, t_pdflat, reward_target = ob.contiguous(), action.contiguous(), t_pdflat.contiguous(), reward_target.contiguous()
        s = torch.empty((T, B, 4), dtype=torch.float32, device=self.device)
        rew = torch.empty((T, B), dtype=torch.float32, device=self.device)
        check(lib().rb_lstm2_loss_grad(self._spec_p, ptr(self.params), ptr(ob), ptr(action), ptr(t_pdflat), ptr(reward_target),
                                       ptr(state.contiguous() if state is not None else None), B, float(keep_prob), int(seed), int(sample_id0),
                                       int(iteration), loss_kind, ptr(s), ptr(rew), None, ptr(self.gradloss), ptr(self._workspace(B)), stream_ptr()))
        return s, rew

    def adam_step(self, grad_scale=1.0):
        self.t += 1
        check(lib().rb_adam_step(ptr(self.params), ptr(self.m), ptr(self.v), ptr(self.gradloss), self.P, self.t, self.lr, self.beta1, self.beta2,
                                 self.eps, grad_scale, stream_ptr()))

    def step(self, ob, action, t_pdflat, reward_target, state=None, keep_prob=1.0, seed=0, sample_id0=0, loss_kind=LOSS_KL_ST, use_graph=True):
        """loss_grad + Adam as one CUDA-graph launch (rb_lstm2_step).  Pass the same (static) input tensors every step -- the graph is re-captured
        whenever a pointer or scalar changes; the dropout iteration and the Adam step count live in a device-side clock.  Returns (s_pdflat, reward)
        in static output buffers."""
        import ctypes as C
        T, B = ob.shape[0], ob.shape[1]
        assert T == self.T and ob.is_contiguous() and action.is_contiguous() and t_pdflat.is_contiguous() and reward_target.is_contiguous()
        if getattr(self, "_ctx", None) is None:
            h = C.c_void_p()
            check(lib().rb_lstm_ctx_create(C.byref(h), self.device.index))
            self._ctx, self._clock_t, self._out_static = h, None, {}
        if self._clock_t != self.t:                       # (re)load the device clock after any non-graph update
            check(lib().rb_lstm_ctx_set_clock(self._ctx, self.t, self.t, stream_ptr()))
        if B not in self._out_static:
            self._out_static[B] = (torch.empty((T, B, 4), dtype=torch.float32, device=self.device), torch.empty((T, B), dtype=torch.float32, device=self.device))
        s_out, rew_out = self._out_static[B]
        check(lib().rb_lstm2_step(self._ctx, self._spec_p, ptr(self.params), ptr(self.m), ptr(self.v), ptr(ob), ptr(action), ptr(t_pdflat), ptr(reward_target),
                                  ptr(state.contiguous() if state is not None else None), B, float(keep_prob), int(seed), int(sample_id0), loss_kind,
                                  ptr(s_out), ptr(rew_out), ptr(self.gradloss), ptr(self._workspace(B)), self.lr, self.beta1, self.beta2, self.eps, 1.0,
                                  1 if use_graph else 0, stream_ptr()))
        self.t += 1
        self._clock_t = self.t
        return s_out, rew_out

    def state_dict(self):
        return dict(kind="lstm2", spec=self.spec.tolist(), params=self.params.cpu(), m=self.m.cpu(), v=self.v.cpu(), t=self.t)

    def load_state_dict(self, sd):
        assert sd["kind"] == "lstm2" and list(sd["spec"]) == self.spec.tolist()
        self.params.copy_(sd["params"]); self.m.copy_(sd["m"]); self.v.copy_(sd["v"]); self.t = int(sd["t"])


def lstm_graph(input_ob, input_action, hidden_combined, net):
    """Reference name (backup/student_rollout.py:130): forward of the two-headed graph -> (s_pdflat, final_state, reward)."""
    s, rew, fin = net.forward(input_ob, input_action, hidden_combined)
    return s, fin, rew
