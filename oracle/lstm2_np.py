"""TEST INFRASTRUCTURE (oracle): float64 numpy restatement of the reference's two-headed LSTM student, /root/reference
src/distilation/backup/student_rollout.py:130-200 (lstm_graph, lstm_loss), :303-328 (placeholders, total loss), with a hand-derived
back-propagation through time.  Cell: tf.contrib.rnn.LSTMCell (TF 1.10): z = [x, m_prev] W + b; i, j, f, o = split(z, 4);
c = sigmoid(f + 1) c_prev + sigmoid(i) tanh(j); m = sigmoid(o) tanh(c).
spec = (units, steps, carry_state, trunk, action_hidden, n_reward_hidden, reward_hidden...):
    SOURCE_SPEC(units, steps)   the checked-in source: `state` is never reassigned inside the unroll (:156) => carry_state = 0, every step
                                starts from the fed state and final_state = the fed state (:191); one hidden reward layer of 64
    TFEVENTS_SPEC               the graph recorded in the reference's tfevents files (tests/golden/graph_facts.json): units 1, 2 steps, state
                                carried, reward head 64-32-64
Parity: pinned to those two sources structurally (tests/test_lstm2_oracle.py checks the parameter shapes against graph_facts.json and the
gradient against central finite differences); TF itself is not installable here, so no recorded activations exist for this graph.
Only tests / smoke may import this module."""
import numpy as np

IN = 13


def SOURCE_SPEC(units=1, steps=2):
    return (units, steps, 0, 128, 64, 1, 64)


TFEVENTS_SPEC = (1, 2, 1, 128, 64, 3, 64, 32, 64)


class Layout:
    def __init__(self, spec):
        self.U, self.T, self.carry, self.D, self.A, self.nR = (int(v) for v in spec[:6])
        self.rh = [int(v) for v in spec[6:6 + self.nR]]
        self.XH, self.G = IN + self.U, 4 * self.U
        self.o_Wl, self.o_bl = 0, self.XH * self.G
        self.head0 = self.o_bl + self.G
        o, self.blocks = 0, []                       # (name, offset of W, offset of b, fan_in, fan_out) inside a step's block, creation order
        def add(name, fi, fo):
            nonlocal o
            self.blocks.append((name, o, o + fi * fo, fi, fo))
            o += fi * fo + fo
        add("trunk", self.U, self.D)
        prev = self.D
        for k, r in enumerate(self.rh):
            add("reward%d" % k, prev, r)
            prev = r
        add("reward_out", prev, 1)
        add("action", self.D, self.A)
        add("pd", self.A, 4)
        self.head_sz = o
        self.P = self.head0 + self.T * self.head_sz

    def spec_array(self):
        a = np.zeros(10, np.int32)
        a[:6] = (self.U, self.T, self.carry, self.D, self.A, self.nR)
        a[6:6 + self.nR] = self.rh
        return a


def param_count(spec):
    return Layout(spec).P


def init_params(spec, seed=0):
    """glorot-uniform kernels, zero biases (tf.layers.dense / LSTMCell get_variable defaults)."""
    L = Layout(spec)
    rng = np.random.default_rng(seed)
    p = np.zeros(L.P, np.float32)
    def glorot(off, fi, fo):
        lim = np.sqrt(6.0 / (fi + fo))
        p[off:off + fi * fo] = rng.uniform(-lim, lim, fi * fo).astype(np.float32)
    glorot(L.o_Wl, L.XH, L.G)
    for t in range(L.T):
        for _, ow, _, fi, fo in L.blocks:
            glorot(L.head0 + t * L.head_sz + ow, fi, fo)
    return p


def _heads(L, p, t):
    base, out = L.head0 + t * L.head_sz, {}
    for name, ow, ob, fi, fo in L.blocks:
        out[name] = (p[base + ow:base + ow + fi * fo].reshape(fi, fo), p[base + ob:base + ob + fo])
    return out


def sig(x):
    return 1.0 / (1.0 + np.exp(-x))


def forward(spec, p, obd, action, state=None):
    """obd [T,B,11] (already dropped out), action [T,B,2], state [2,B,U] or None -> pdflat [T,B,4], reward [T,B], final state [2,B,U], cache."""
    L = Layout(spec)
    p = np.asarray(p, np.float64)
    Wl, bl = p[L.o_Wl:L.o_bl].reshape(L.XH, L.G), p[L.o_bl:L.head0]
    B, U = obd.shape[1], L.U
    c0 = np.zeros((B, U)) if state is None else np.asarray(state[0], np.float64)
    m0 = np.zeros((B, U)) if state is None else np.asarray(state[1], np.float64)
    c, m = c0, m0
    s, rew, cache = np.zeros((L.T, B, 4)), np.zeros((L.T, B)), []
    for t in range(L.T):
        c_prev, m_prev = (c, m) if L.carry else (c0, m0)
        xh = np.concatenate([np.asarray(obd[t], np.float64), np.asarray(action[t], np.float64), m_prev], -1)
        z = xh @ Wl + bl
        i, j, f, o = sig(z[:, :U]), np.tanh(z[:, U:2 * U]), sig(z[:, 2 * U:3 * U] + 1.0), sig(z[:, 3 * U:])
        c = f * c_prev + i * j
        m = o * np.tanh(c)
        H = _heads(L, p, t)
        trunk = np.tanh(m @ H["trunk"][0] + H["trunk"][1])
        r_acts, a = [], trunk
        for k in range(L.nR):
            a = np.tanh(a @ H["reward%d" % k][0] + H["reward%d" % k][1])
            r_acts.append(a)
        rew[t] = (a @ H["reward_out"][0] + H["reward_out"][1])[:, 0]
        a1 = np.tanh(trunk @ H["action"][0] + H["action"][1])
        s[t] = a1 @ H["pd"][0] + H["pd"][1]
        cache.append((xh, i, j, f, o, c_prev, c, m, trunk, r_acts, a1))
    final = np.stack([c, m]) if L.carry else np.stack([c0, m0])
    return s, rew, final, cache


def kl(s, t):
    """lstm_loss (:196-200): KL(student || teacher) summed over everything, and d/ds."""
    s, t = np.asarray(s, np.float64), np.asarray(t, np.float64)
    ms, ls, mt, lt = s[..., :2], s[..., 2:], t[..., :2], t[..., 2:]
    vs, vt = np.exp(2 * ls), np.exp(2 * lt)
    loss = (lt - ls + (vs + (ms - mt) ** 2) / (2 * vt) - 0.5).sum()
    return loss, np.concatenate([(ms - mt) / vt, vs / vt - 1.0], -1)


def loss_grad(spec, p, obd, action, t_pd, reward_target, state=None):
    """-> (pdflat, reward, (total, kl, reward part), flat gradient[P])."""
    L = Layout(spec)
    p = np.asarray(p, np.float64)
    Wl = p[L.o_Wl:L.o_bl].reshape(L.XH, L.G)
    s, rew, final, cache = forward(spec, p, obd, action, state)
    lk, ds = kl(s, t_pd)
    e = rew - np.asarray(reward_target, np.float64)
    lr = (e ** 2).sum()
    g = np.zeros(L.P)
    B, U = obd.shape[1], L.U
    dm_next, dc_next = np.zeros((B, U)), np.zeros((B, U))
    gWl, gbl = np.zeros((L.XH, L.G)), np.zeros(L.G)
    for t in range(L.T - 1, -1, -1):
        xh, i, j, f, o, c_prev, c, m, trunk, r_acts, a1 = cache[t]
        H = _heads(L, p, t)
        base = L.head0 + t * L.head_sz
        off = {name: (base + ow, base + ob) for name, ow, ob, _, _ in L.blocks}
        def put(name, inp, d):
            ow, ob = off[name]
            g[ow:ow + inp.shape[1] * d.shape[1]] = (inp.T @ d).ravel()
            g[ob:ob + d.shape[1]] = d.sum(0)
        # action head
        d = ds[t]
        put("pd", a1, d)
        d = (d @ H["pd"][0].T) * (1 - a1 ** 2)
        put("action", trunk, d)
        dtrunk = d @ H["action"][0].T
        # reward head
        d = (2.0 * e[t])[:, None]
        put("reward_out", r_acts[-1], d)
        d = d @ H["reward_out"][0].T
        for k in range(L.nR - 1, -1, -1):
            d = d * (1 - r_acts[k] ** 2)
            put("reward%d" % k, r_acts[k - 1] if k > 0 else trunk, d)
            d = d @ H["reward%d" % k][0].T
        dtrunk = (dtrunk + d) * (1 - trunk ** 2)
        put("trunk", m, dtrunk)
        dm = dtrunk @ H["trunk"][0].T + (dm_next if L.carry else 0.0)
        tc = np.tanh(c)
        dct = (dc_next if L.carry else 0.0) + dm * o * (1 - tc ** 2)
        dz = np.concatenate([dct * j * i * (1 - i), dct * i * (1 - j ** 2), dct * c_prev * f * (1 - f), dm * tc * o * (1 - o)], -1)
        dc_next = dct * f
        gWl += xh.T @ dz
        gbl += dz.sum(0)
        dm_next = (dz @ Wl.T)[:, IN:]
    g[L.o_Wl:L.o_bl], g[L.o_bl:L.head0] = gWl.ravel(), gbl
    return s, rew, (lk + lr, lk, lr), g
