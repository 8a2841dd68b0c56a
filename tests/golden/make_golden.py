#!/usr/bin/env python
"""Generate tests/golden/reacher_fixture.npz from the reference's own recorded MuJoCo data.

Source: /root/reference/src/distilation/tests/data/dataset.json (gzip'd json_tricks page written by
dataset.py:31-34; the fixture tests/dataset_unit_test.py:114 points at).  25 episodes x 50 steps of real
gym-0.10.5 / MuJoCo-1.50 Reacher-v2 transitions with teacher / student pdflat.  This script only runs in the
build container (the reference tree does not travel to the GPU box); its output is committed.
"""
import gzip
import json
import sys

import numpy as np

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/distilation/tests/data/dataset.json"
DST = sys.argv[2] if len(sys.argv) > 2 else __file__.rsplit("/", 1)[0] + "/reacher_fixture.npz"

d = json.loads(gzip.decompress(open(SRC, "rb").read()))
E, T = len(d), len(d[0])
ob = np.array([[s["ob"] for s in ep] for ep in d], dtype=np.float64)
rew = np.array([[float(np.ravel(s["rew"])[0]) for s in ep] for ep in d], dtype=np.float64)
t = np.array([[s["t"] for s in ep] for ep in d], dtype=np.float64)
s_ = np.array([[s["s"] for s in ep] for ep in d], dtype=np.float64)
prev = np.array([[s["prev"] for s in ep] for ep in d], dtype=np.float64)
with_s = np.array([[1 if s["with"] == "s" else 0 for s in ep] for ep in d], dtype=np.uint8)
assert ob.shape == (E, T, 11) and t.shape == (E, T, 4)
np.savez_compressed(DST, ob=ob, rew=rew, t=t, s=s_, prev=prev, with_s=with_s)
print("wrote", DST, ob.shape, "student-stepped episodes:", np.where(with_s.all(1))[0])


def make_reference_page(src="/root/reference/src/distilation/tests/data/dataset.json", dst=None, episodes=2):
    """tests/golden/dataset_page_ref.json: the first `episodes` episodes of the reference's own page file, re-compressed unchanged
    (gzip of the json_tricks primitives JSON) -- pins the on-disk page format (dataset.py:31-48)."""
    import gzip
    import json
    import os
    dst = dst or os.path.join(os.path.dirname(os.path.abspath(__file__)), "dataset_page_ref.json")
    d = json.loads(gzip.decompress(open(src, "rb").read()))
    open(dst, "wb").write(gzip.compress(json.dumps(d[:episodes]).encode()))
