"""TEST INFRASTRUCTURE (oracle): float64 numpy restatement of the reference's LSTM student, /root/reference
src/distilation/student_nn.py:21-49 (graph) + loss.py:3-13 (KL), with a hand-derived back-propagation through time.
tf.contrib.rnn.LSTMCell (TF 1.10, requirement.txt:62): z = [x, m_prev] W + b; i, j, f, o = split(z, 4);
c = sigmoid(f + forget_bias(=1)) c_prev + sigmoid(i) tanh(j); m = sigmoid(o) tanh(c); state tuple (c, m).
Flat parameter layout = include/reacher_b200.h (rb_lstm_param_count).  Only tests / smoke may import this module."""
import numpy as np

T, U, G, X, E, XH = 10, 200, 800, 43, 32, 243
HD = (200, 64, 128, 64, 32, 4)
L_WE, L_BE, L_WL = 0, 128, 160
L_BL = L_WL + XH * G
L_HEAD0 = L_BL + G
L_HEAD_SZ = sum(HD[i] * HD[i + 1] + HD[i + 1] for i in range(5))
P = L_HEAD0 + T * L_HEAD_SZ


def param_count():
    return P


def init_params(seed=0):
    """glorot-uniform kernels, zero biases (tf.layers.dense / LSTMCell get_variable defaults)."""
    rng = np.random.default_rng(seed)
    p = np.zeros(P, np.float32)
    def glorot(off, fi, fo):
        lim = np.sqrt(6.0 / (fi + fo))
        p[off:off + fi * fo] = rng.uniform(-lim, lim, fi * fo).astype(np.float32)
    glorot(L_WE, 4, E)
    glorot(L_WL, XH, G)
    for t in range(T):
        o = L_HEAD0 + t * L_HEAD_SZ
        for l in range(5):
            glorot(o, HD[l], HD[l + 1])
            o += HD[l] * HD[l + 1] + HD[l + 1]
    return p


def _views(p):
    p = np.asarray(p, np.float64)
    We, be = p[L_WE:L_BE].reshape(4, E), p[L_BE:L_WL]
    Wl, bl = p[L_WL:L_BL].reshape(XH, G), p[L_BL:L_HEAD0]
    heads = []
    for t in range(T):
        o, hs = L_HEAD0 + t * L_HEAD_SZ, []
        for l in range(5):
            W = p[o:o + HD[l] * HD[l + 1]].reshape(HD[l], HD[l + 1]); o += HD[l] * HD[l + 1]
            b = p[o:o + HD[l + 1]]; o += HD[l + 1]
            hs.append((W, b))
        heads.append(hs)
    return We, be, Wl, bl, heads


def sig(x):
    return 1.0 / (1.0 + np.exp(-x))


def forward(p, obd, prev_pd, state=None):
    """obd [T,B,11] (already dropped out), prev_pd [T,B,4], state [2,B,200] or None -> s [T,B,4], final state, cache."""
    We, be, Wl, bl, heads = _views(p)
    B = obd.shape[1]
    c = np.zeros((B, U)) if state is None else np.asarray(state[0], np.float64)
    m = np.zeros((B, U)) if state is None else np.asarray(state[1], np.float64)
    emb = np.asarray(prev_pd, np.float64) @ We + be
    cache, s = [], np.zeros((T, B, 4))
    for t in range(T):
        xh = np.concatenate([np.asarray(obd[t], np.float64), emb[t], m], -1)
        z = xh @ Wl + bl
        i, j, f, o = sig(z[:, :U]), np.tanh(z[:, U:2 * U]), sig(z[:, 2 * U:3 * U] + 1.0), sig(z[:, 3 * U:])
        c_prev = c
        c = f * c_prev + i * j
        m = o * np.tanh(c)
        acts, a = [m], m
        for l in range(5):
            W, b = heads[t][l]
            a = a @ W + b
            if l < 4:
                a = np.tanh(a)
            acts.append(a)
        s[t] = a
        cache.append((xh, i, j, f, o, c_prev, c, acts))
    return s, np.stack([c, m]), cache


def kl(s, t, reverse=False):
    s, t = np.asarray(s, np.float64), np.asarray(t, np.float64)
    ms, ls, mt, lt = s[..., :2], s[..., 2:], t[..., :2], t[..., 2:]
    vs, vt = np.exp(2 * ls), np.exp(2 * lt)
    if not reverse:
        loss = (lt - ls + (vs + (ms - mt) ** 2) / (2 * vt) - 0.5).sum()
        ds = np.concatenate([(ms - mt) / vt, vs / vt - 1.0], -1)
    else:
        loss = (ls - lt + (vt + (ms - mt) ** 2) / (2 * vs) - 0.5).sum()
        ds = np.concatenate([(ms - mt) / vs, 1.0 - (vt + (ms - mt) ** 2) / vs], -1)
    return loss, ds


def loss_grad(p, obd, prev_pd, t_pd, state=None, reverse=False):
    """-> (s, loss, flat gradient[P]) by back-propagation through time."""
    We, be, Wl, bl, heads = _views(p)
    s, final, cache = forward(p, obd, prev_pd, state)
    loss, ds = kl(s, t_pd, reverse)
    g = np.zeros(P)
    B = obd.shape[1]
    dm_next, dc_next = np.zeros((B, U)), np.zeros((B, U))
    gWl, gbl, gWe, gbe = np.zeros((XH, G)), np.zeros(G), np.zeros((4, E)), np.zeros(E)
    for t in range(T - 1, -1, -1):
        xh, i, j, f, o, c_prev, c, acts = cache[t]
        d = ds[t]
        off = L_HEAD0 + t * L_HEAD_SZ
        offs = [off + sum(HD[k] * HD[k + 1] + HD[k + 1] for k in range(l)) for l in range(5)]
        for l in range(4, -1, -1):
            W, b = heads[t][l]
            g[offs[l]:offs[l] + W.size] = (acts[l].T @ d).ravel()
            g[offs[l] + W.size:offs[l] + W.size + b.size] = d.sum(0)
            d = d @ W.T
            if l > 0:
                d = d * (1.0 - acts[l] ** 2)
        dm = d + dm_next
        tc = np.tanh(c)
        dct = dc_next + dm * o * (1.0 - tc ** 2)
        dz = np.concatenate([dct * j * i * (1 - i), dct * i * (1 - j ** 2), dct * c_prev * f * (1 - f), dm * tc * o * (1 - o)], -1)
        dc_next = dct * f
        gWl += xh.T @ dz
        gbl += dz.sum(0)
        dxh = dz @ Wl.T
        dm_next = dxh[:, X:]
        demb = dxh[:, 11:X]
        gWe += np.asarray(prev_pd[t], np.float64).T @ demb
        gbe += demb.sum(0)
    g[L_WE:L_BE], g[L_BE:L_WL], g[L_WL:L_BL], g[L_BL:L_HEAD0] = gWe.ravel(), gbe, gWl.ravel(), gbl
    return s, loss, g
