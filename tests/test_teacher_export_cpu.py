"""teacher.policy_params_from_named: the variables of the reference's `teacher.ckpt` (names from its tfevents graph) packed
into the flat C-ABI layout -- checked on the host against a direct numpy evaluation of the named graph and the oracle's flat-layout policy."""
import numpy as np
import pytest

from oracle import nn_np as NN


def _named(rng, nout=2, suffix=""):
    count = 12345.0
    x = rng.standard_normal((int(count), 11)) * (0.02 + rng.random(11)) + rng.standard_normal(11)   # one dim nearly constant -> variance floor
    x[:, 10] = 0.0                                                                                   # Reacher obs[10] is always 0
    v = {"pi/obfilter/runningsum": x.sum(0), "pi/obfilter/runningsumsq": (x * x).sum(0), "pi/obfilter/count": np.array(count),
         "pi/pol/fc1/w": rng.standard_normal((11, 64)) * 0.3, "pi/pol/fc1/b": rng.standard_normal(64) * 0.1,
         "pi/pol/fc2/w": rng.standard_normal((64, 64)) * 0.2, "pi/pol/fc2/b": rng.standard_normal(64) * 0.1,
         "pi/pol/final/w": rng.standard_normal((64, nout)) * 0.2, "pi/pol/final/b": rng.standard_normal(nout) * 0.1,
         "pi/pol/logstd": np.array([[-3.29, -3.36]]),
         "pi/vf/fc1/w": rng.standard_normal((11, 64))}                                             # the unused value tower is ignored
    return {k + suffix: a for k, a in v.items()}


def _graph(v, ob):
    mu = v["pi/obfilter/runningsum"] / v["pi/obfilter/count"]
    sd = np.sqrt(np.maximum(v["pi/obfilter/runningsumsq"] / v["pi/obfilter/count"] - mu * mu, 1e-2))
    z = np.clip((ob - mu) / sd, -5.0, 5.0)
    h = np.tanh(z @ v["pi/pol/fc1/w"] + v["pi/pol/fc1/b"])
    h = np.tanh(h @ v["pi/pol/fc2/w"] + v["pi/pol/fc2/b"])
    mean = h @ v["pi/pol/final/w"] + v["pi/pol/final/b"]
    return np.concatenate([mean, 0.0 * mean + v["pi/pol/logstd"]], -1)


@pytest.mark.parametrize("suffix", ["", ":0"])
def test_named_checkpoint_variables_give_the_same_policy(suffix):
    from reacherdistilation_b200.teacher import policy_params_from_named
    rng = np.random.default_rng(7)
    v = _named(rng, suffix=suffix)
    p = policy_params_from_named(v)
    assert p.dtype == np.float32 and p.size == 5058 + 22 + 2
    plain = {k[:-2] if k.endswith(":0") else k: a for k, a in v.items()}
    assert p[21] == np.float32(0.1)                                  # obs[10] == 0 always: the variance floor 0.01 -> std 0.1
    ob = (rng.standard_normal((200, 11)) * 2).astype(np.float32)
    ref = _graph(plain, ob.astype(np.float64))
    assert np.abs(NN.policy_fwd(ob, p) - ref).max() <= 2e-6          # fp32 parameters vs the float64 named graph


def test_kernel_bias_aliases_and_errors():
    from reacherdistilation_b200.teacher import policy_params_from_named
    rng = np.random.default_rng(8)
    v = _named(rng)
    alias = {k.replace("/w", "/kernel").replace("/b", "/bias") if "/pol/f" in k else k: a for k, a in v.items()}
    assert np.array_equal(policy_params_from_named(alias), policy_params_from_named(v))
    bad = dict(v); del bad["pi/pol/fc2/b"]
    with pytest.raises(KeyError):
        policy_params_from_named(bad)
    bad = dict(v); bad["pi/pol/fc1/w"] = np.zeros((64, 11))
    with pytest.raises(ValueError):
        policy_params_from_named(bad)
