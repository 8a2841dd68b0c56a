"""DatasetStore (`dataset.py:14-65`): the reference's page files -- gzip(json) written by json_tricks(primitives=True, compression=True) --
on the host only (no GPU): file naming, numeric page order, refusal to overwrite, and a page written by the REFERENCE itself
(tests/golden/dataset_page_ref.json, cut from its tests/data/dataset.json by tests/golden/make_golden.py)."""
import gzip
import json
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dataset_page_ref.json")


def _episode(seed, steps=50):
    rng = np.random.default_rng(seed)
    return [{"ob": rng.standard_normal(11).tolist(), "rew": [float(rng.standard_normal())], "t": rng.standard_normal(4).tolist(),
             "s": [0.0] * 4, "with": "t", "prev": [0.0] * 4, "prew": [0.0]} for _ in range(steps)]


def test_store_load_round_trip_and_naming(tmp_path):
    from reacherdistilation_b200.dataset import DatasetStore
    st = DatasetStore(str(tmp_path / "pages"))
    assert st.pages == [] and st.rand_pages(3) is None
    written = [st.store([_episode(10 * p + e) for e in range(3)]) for p in range(12)]
    assert [os.path.basename(p) for p in written] == ["dataset_%d.json" % p for p in range(12)]       # dataset.py:28-29 file names
    back = DatasetStore.load(written[7])
    assert back == [_episode(70 + e) for e in range(3)]                                                 # lossless (plain JSON floats)
    with open(written[7], "rb") as fh:                                                                  # the format the reference's reader expects
        assert json.loads(gzip.decompress(fh.read())) == back
    again = DatasetStore(str(tmp_path / "pages"))                                                      # re-open: pages in NUMERIC order (10 after 9)
    assert again.pages == written
    picks = again.rand_pages(5, np.random.default_rng(1))
    assert len(picks) == 5 and len(set(picks)) == 5 and set(picks) <= set(written)
    assert len(again.rand_pages(99)) == 12


def test_store_refuses_to_overwrite(tmp_path):
    from reacherdistilation_b200.dataset import DatasetStore
    st = DatasetStore(str(tmp_path))
    st.store([_episode(0)])
    open(st.get_full_path(1), "wb").close()                     # someone else created the next page meanwhile
    with pytest.raises(FileExistsError):                        # dataset.py:55-61
        st.store([_episode(1)])


def test_reads_a_page_written_by_the_reference():
    from reacherdistilation_b200.dataset import DatasetStore
    eps = DatasetStore.load(GOLDEN)
    assert len(eps) == 2 and all(len(ep) == 50 for ep in eps)
    for ep in eps:
        assert set(ep[0]) >= {"ob", "rew", "t", "s", "with", "prev"} and len(ep[0]["ob"]) == 11 and len(ep[0]["t"]) == 4
        assert np.allclose(ep[0]["prev"], 0.0)                                         # zeros at k = 0 (dataset.py:118-143)
        for k in range(1, 50):
            if ep[k]["with"] == "t":
                assert np.allclose(ep[k]["prev"], ep[k - 1]["t"])                       # `prev` = teacher pdflat of the previous record
        ob = np.asarray([r["ob"] for r in ep])
        assert np.allclose(ob[:, 0] ** 2 + ob[:, 2] ** 2, 1.0, atol=1e-12)             # cos^2 + sin^2 of joint 0
