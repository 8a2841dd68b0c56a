"""Device Dataset (rb_dataset_*) vs the CPU restatement of the reference Dataset: pure data movement, so everything is bit-exact."""
import numpy as np
import pytest
import torch

from oracle.dataset_np import DatasetOracle

pytestmark = pytest.mark.gpu


def _fill(n, gens, G, seed, partial=0):
    from reacherdistilation_b200.dataset import Dataset
    rng = np.random.default_rng(seed)
    dev, ref = Dataset(num_envs=n, generations=G, seed=seed), DatasetOracle(num_envs=n, generations=G, seed=seed)
    def step(with_):
        ob, rw = rng.standard_normal((n, 11)).astype(np.float32), rng.standard_normal(n).astype(np.float32)
        t, s = rng.standard_normal((n, 4)).astype(np.float32), rng.standard_normal((n, 4)).astype(np.float32)
        dev.write(torch.from_numpy(ob).cuda(), torch.from_numpy(rw).cuda(), torch.from_numpy(t).cuda(), torch.from_numpy(s).cuda(), with_)
        ref.write(ob, rw, t, s, with_)
    for g in range(gens):
        for k in range(50):
            step("t" if g == 0 else "s")
        dev.flush(); ref.flush()
    for k in range(partial):
        step("s")
    return dev, ref, rng


@pytest.mark.parametrize("n,gens,G,partial", [(1, 3, 8, 0), (7, 2, 2, 0), (64, 5, 3, 0), (64, 5, 3, 4), (300, 1, 4, 49)])
def test_training_batches_match_oracle(n, gens, G, partial):
    dev, ref, _ = _fill(n, gens, G, seed=11, partial=partial)
    assert dev.num_episodes() == ref.num_total_episodes == gens * n
    for draw in range(6):
        for B, T in ((20, 10), (33, 1), (5, 50)):
            ob, t, prev, prew, eps, start = dev.training_batch(B, T, draw=draw, with_indices=True)
            r = ref.training_batch(draw, B, T)
            assert np.array_equal(eps.cpu().numpy(), r[4]) and int(start.item()) == r[5]
            for a, b in zip((ob, t, prev, prew), r[:4]):
                assert np.array_equal(a.cpu().numpy().astype(np.float64), b)
    dev.close()


@pytest.mark.parametrize("length", [0, 1, 7, 9, 10, 15, 49])
def test_test_batch_matches_oracle(length):
    n = 33
    dev, ref, rng = _fill(n, 1, 4, seed=3, partial=length)
    ob = rng.standard_normal((n, 11)).astype(np.float32)
    o, p, w = dev.test_batch(torch.from_numpy(ob).cuda())
    ro, rp, rw = ref.test_batch(ob)
    assert dev.last_step() == length - 1
    assert np.array_equal(o.cpu().numpy().astype(np.float64), ro) and np.array_equal(p.cpu().numpy().astype(np.float64), rp)
    assert np.array_equal(w.cpu().numpy().astype(np.float64), rw)
    dev.close()


def test_single_env_reference_shape_and_errors():
    """num_envs == 1: zero batches [T, B, .] with the window in the LAST batch row (dataset.py:239,271,288); episode-length guards."""
    from reacherdistilation_b200 import ReacherB200Error
    from reacherdistilation_b200.dataset import Dataset
    ds = Dataset(num_envs=1, generations=2)
    with pytest.raises(ReacherB200Error):
        ds.training_batch()                                   # nothing flushed yet
    for k in range(12):
        ds.write(np.full((1, 11), float(k)), [0.5 * k], np.full((1, 4), 10.0 + k))
    o, p, w = ds.test_batch(np.full((1, 11), -10.0))
    assert o.shape == (10, 20, 11) and p.shape == (10, 20, 4) and w.shape == (10, 20, 1)
    assert float(o[:, :19].abs().max()) == 0.0 and float(p[:, :19].abs().max()) == 0.0
    assert o[:9, 19, 0].tolist() == [3.0, 4.0, 5.0, 6.0, 7.0, 8.0, 9.0, 10.0, 11.0] and o[9, 19, 0].item() == -10.0
    assert p[:, 19, 0].tolist() == [12.0, 13.0, 14.0, 15.0, 16.0, 17.0, 18.0, 19.0, 20.0, 21.0]      # t of records 2..11
    with pytest.raises(ReacherB200Error):
        ds.flush()                                            # only complete (50-record) episodes are flushed
    for k in range(38):
        ds.write(np.zeros((1, 11)))
    ds.flush()
    with pytest.raises(ValueError):
        ds.dump()                                             # no dir_path given
    assert ds.num_episodes() == 1
    ds.close()


def test_page_store_round_trip_and_reference_page(tmp_path):
    """Dataset.dump -> dataset_<k>.json pages (gzip JSON, MAX_CAPACITY episodes each) -> switch / load back; and a page written by
    the REFERENCE itself (tests/golden/dataset_page_ref.json, 2 recorded MuJoCo episodes) loads into the device ring."""
    import os
    from reacherdistilation_b200.dataset import Dataset, DatasetStore
    n = 25
    dev, ref, _ = _fill(n, 2, 4, seed=5)
    dev.dstore = DatasetStore(str(tmp_path))
    pages = dev.dump()
    assert [os.path.basename(p) for p in pages] == ["dataset_%d.json" % i for i in range(6)]          # 50 episodes / 10 per page, per generation 3
    assert dev.dump() == []                                                                         # nothing new
    eps = [ep for p in dev.pages() for ep in dev.switch(p)]
    assert len(eps) == 2 * n and all(len(ep) == 50 for ep in eps)
    for e, ep in enumerate(eps):
        r = ref.data_in_memory[e]
        for k in (0, 1, 17, 49):
            assert np.array_equal(np.float32(ep[k]["ob"]), np.float32(r[k]["ob"])) and np.array_equal(np.float32(ep[k]["prev"]), np.float32(r[k]["prev"]))
            assert ep[k]["with"] == r[k]["with"] and np.float32(np.ravel(ep[k]["rew"])[0]) == np.float32(r[k]["rew"])
    dev.close()
    # a page in the reference's own format
    page = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dataset_page_ref.json")
    one = Dataset(num_envs=1, generations=4)
    assert one.load_page_into_ring(page) == 2 and one.num_episodes() == 2
    raw = DatasetStore.load(page)
    ob, t, prev, prew, eps_idx, start = one.training_batch(4, 10, draw=0, with_indices=True)
    s0 = int(start.item())
    for b, e in enumerate(eps_idx.cpu().numpy()):
        assert np.array_equal(ob[:, b].cpu().numpy(), np.float32([raw[e][k]["ob"] for k in range(s0, s0 + 10)]))
        assert np.array_equal(prev[:, b].cpu().numpy(), np.float32([raw[e][k]["prev"] for k in range(s0, s0 + 10)]))   # recorded prev == t[k-1]
    one.close()
