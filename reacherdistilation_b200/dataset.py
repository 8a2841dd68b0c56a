"""Dataset -- the reference's rollout buffer (/root/reference src/distilation/dataset.py:72-296) on the device.

Same method names: write / flush / training_batches / test_batch / num_episodes / last_step.  N lock-step envs append one
record each per call; sampling uses Philox (seed, draw counter) instead of Python's `random`.  DatasetStore keeps the reference's
on-disk page format (dataset.py:14-65): `dataset_<k>.json` = gzip(JSON list[episode][step]{ob, rew, t, s, with, prev, prew}), at most
MAX_CAPACITY episodes per page, so extract_reward.py / plot.py keep working on exported rollouts.
"""
import ctypes as C
import gzip
import json
import os
import re

import numpy as np
import torch

from ._lib import check, lib, ptr, stream_ptr
from .config import EPISODE_STEPS, LSTM_BATCH_SIZE, MAX_CAPACITY, STEPS_UNROLLED, TRAINING_EPOCHS

__all__ = ["Dataset", "DatasetStore"]


class DatasetStore:
    """Page files of the reference (dataset.py:14-65).  json_tricks(primitives=True, compression=True) == gzip(json) of plain lists."""

    def __init__(self, dir_path):
        self.dir_path = dir_path
        os.makedirs(dir_path, exist_ok=True)
        self.pages = self.collect_pages(dir_path)

    def get_full_path(self, page):
        return os.path.join(self.dir_path, "dataset_%d.json" % page)

    def collect_pages(self, dir_path):
        fs = [f for f in os.listdir(dir_path) if re.fullmatch(r"dataset_(\d+)\.json", f)]
        return [os.path.join(dir_path, f) for f in sorted(fs, key=lambda f: int(re.search(r"(\d+)", f).group(1)))]

    def store(self, episodes):
        """Write `episodes` (list[episode][step] of dicts) as ONE new page; refuses to overwrite (dataset.py:55-61)."""
        path = self.get_full_path(len(self.pages))
        if os.path.exists(path):
            raise FileExistsError("current page already exists. will not overwrite")
        with open(path, "wb") as fh:
            fh.write(gzip.compress(json.dumps(episodes).encode()))
        self.pages.append(path)
        return path

    @staticmethod
    def load(page):
        with open(page, "rb") as fh:
            return json.loads(gzip.decompress(fh.read()))

    def rand_pages(self, num_pages, rng=None):
        if not self.pages:
            return None
        rng = rng or np.random.default_rng(0)
        return [self.pages[i] for i in rng.choice(len(self.pages), size=min(num_pages, len(self.pages)), replace=False)]


class Dataset:
    def __init__(self, dir_path=None, num_envs=1, generations=64, device=0, seed=0):
        self.dir_path, self.num_envs, self.seed = dir_path, int(num_envs), int(seed)
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        h = C.c_void_p()
        check(lib().rb_dataset_create(C.byref(h), self.num_envs, int(generations), self.device.index))
        self._h = h
        self._draw = 0
        self.dstore = DatasetStore(dir_path) if dir_path else None
        self.data_in_memory = []          # host copy of the page last loaded with switch() (what extract_reward.py iterates)
        self._dumped = 0                  # generations already written to pages

    # ---- recording -----------------------------------------------------------------------------------------
    def _dev(self, a, width):
        if a is None:
            return None
        t = torch.as_tensor(a, dtype=torch.float32).to(self.device).reshape(self.num_envs, width) if width else \
            torch.as_tensor(a, dtype=torch.float32).to(self.device).reshape(self.num_envs)
        return t.contiguous()

    def write(self, ob, reward=None, t_pdflat=None, s_pdflat=None, stepped_with="t"):
        """dataset.py:118-143.  ob [N,11], reward [N], pdflats [N,4] (device tensors or array-likes)."""
        args = (ob, reward, t_pdflat, s_pdflat)               # as given: host arrays become temporaries on the device below
        ob, rw, t, s = self._dev(ob, 11), self._dev(reward, 0), self._dev(t_pdflat, 4), self._dev(s_pdflat, 4)
        check(lib().rb_dataset_write(self._h, ptr(ob), ptr(rw), ptr(t), ptr(s), 0 if stepped_with == "t" else 1, stream_ptr()))
        if any(a is not None and not torch.is_tensor(a) for a in args):
            torch.cuda.current_stream().synchronize()        # temporaries made from host data must outlive the copy

    def flush(self):
        check(lib().rb_dataset_flush(self._h))

    def num_episodes(self):
        return int(lib().rb_dataset_num_episodes(self._h))

    def last_step(self):
        return int(lib().rb_dataset_episode_len(self._h)) - 1

    # ---- batches -------------------------------------------------------------------------------------------
    def training_batch(self, batch_size=LSTM_BATCH_SIZE, steps=STEPS_UNROLLED, draw=None, with_indices=False, out=None):
        """One batch of dataset.py:184-194: (ob[T,B,11], t[T,B,4], prev[T,B,4], prew[T,B,1]) as device tensors.
        out: the four tensors to fill (static buffers keep a captured CUDA graph of the optimiser step valid from batch to batch)."""
        if draw is None:
            draw, self._draw = self._draw, self._draw + 1
        dev = self.device
        if out is not None:
            ob, t, prev, prew = out
        else:
            ob, t = torch.empty((steps, batch_size, 11), device=dev), torch.empty((steps, batch_size, 4), device=dev)
            prev, prew = torch.empty((steps, batch_size, 4), device=dev), torch.empty((steps, batch_size, 1), device=dev)
        eps = torch.empty(batch_size, dtype=torch.int32, device=dev) if with_indices else None
        start = torch.empty(1, dtype=torch.int32, device=dev) if with_indices else None
        check(lib().rb_dataset_training_batch(self._h, self.seed, int(draw), batch_size, steps, ptr(ob), ptr(t), ptr(prev), ptr(prew), ptr(eps),
                                              ptr(start), stream_ptr()))
        return (ob, t, prev, prew, eps, start) if with_indices else (ob, t, prev, prew)

    def training_batches(self, batch_size=LSTM_BATCH_SIZE, steps=STEPS_UNROLLED):
        for _ in range(TRAINING_EPOCHS):
            yield self.training_batch(batch_size, steps)

    def test_batch(self, ob, steps=STEPS_UNROLLED, batch_size=LSTM_BATCH_SIZE):
        """dataset.py:213-290.  num_envs == 1: the reference's shape -- zero batches [T,B,.] with the window in the LAST batch row.
        num_envs > 1: [T,N,.], column e = window of env e."""
        dev, n = self.device, self.num_envs
        ob = self._dev(ob, 11)
        o, p, r = torch.empty((steps, n, 11), device=dev), torch.empty((steps, n, 4), device=dev), torch.empty((steps, n, 1), device=dev)
        check(lib().rb_dataset_test_batch(self._h, ptr(ob), steps, ptr(o), ptr(p), ptr(r), stream_ptr()))
        if n > 1:
            return o, p, r
        out = [torch.zeros((steps, batch_size, w), device=dev) for w in (11, 4, 1)]
        for dst, src in zip(out, (o, p, r)):
            dst[:, batch_size - 1, :] = src[:, 0, :]
        return tuple(out)

    # ---- page store (dataset.py:80-116) -----------------------------------------------------------------------
    def export_generation(self, generation):
        """One flushed generation as host episodes: list[N] of list[50] of dicts in the reference's record format."""
        n = self.num_envs
        ob, rew = np.empty((EPISODE_STEPS, n, 11), np.float32), np.empty((EPISODE_STEPS, n), np.float32)
        t, s = np.empty((EPISODE_STEPS, n, 4), np.float32), np.empty((EPISODE_STEPS, n, 4), np.float32)
        w = np.empty((EPISODE_STEPS, n), np.uint8)
        check(lib().rb_dataset_export_host(self._h, int(generation), ob.ctypes.data, rew.ctypes.data, t.ctypes.data, s.ctypes.data, w.ctypes.data))
        eps = []
        for e in range(n):
            ep = []
            for k in range(EPISODE_STEPS):
                ep.append({"ob": ob[k, e].astype(float).tolist(), "rew": [float(rew[k, e])], "t": t[k, e].astype(float).tolist(),
                           "s": s[k, e].astype(float).tolist(), "with": "s" if w[k, e] else "t",
                           "prev": t[k - 1, e].astype(float).tolist() if k else [0.0] * 4, "prew": [float(rew[k - 1, e])] if k else [0]})
            eps.append(ep)
        return eps

    def dump(self):
        """Write every flushed generation not yet on disk as pages of MAX_CAPACITY episodes (dataset.py:80-85, :31-40)."""
        if self.dstore is None:
            raise ValueError("Dataset was created without dir_path")
        written = []
        gens = int(lib().rb_dataset_generations(self._h))
        for g in range(self._dumped, gens):
            eps = self.export_generation(g)
            for i in range(0, len(eps), MAX_CAPACITY):
                written.append(self.dstore.store(eps[i:i + MAX_CAPACITY]))
        self._dumped = gens
        return written

    def pages(self):
        return tuple(self.dstore.pages) if self.dstore else ()

    def switch(self, page):
        """Load a page into data_in_memory (host), as the reference's analysis scripts expect (extract_reward.py:25-31)."""
        self.data_in_memory = DatasetStore.load(page)
        return self.data_in_memory

    def load_page_into_ring(self, page):
        """Replay a page (episodes of 50 records) into the device ring, num_envs episodes at a time; returns episodes loaded."""
        eps = [ep for ep in DatasetStore.load(page) if len(ep) == EPISODE_STEPS]
        n, loaded = self.num_envs, 0
        for i in range(0, len(eps) - n + 1, n):
            grp = eps[i:i + n]
            for k in range(EPISODE_STEPS):
                self.write(np.array([ep[k]["ob"] for ep in grp], np.float32), np.array([np.ravel(ep[k]["rew"])[0] for ep in grp], np.float32),
                           np.array([ep[k]["t"] for ep in grp], np.float32), np.array([ep[k]["s"] for ep in grp], np.float32), grp[0][k]["with"])
            self.flush()
            loaded += n
        return loaded

    # ---- exact resume ------------------------------------------------------------------------------------------
    def state_dict(self):
        """The whole ring + cursor + sampling counter (host tensors): a restored loop draws the same windows."""
        rows = int(lib().rb_dataset_ring_rows(self._h))
        ob, rew = np.empty((rows, 11), np.float32), np.empty(rows, np.float32)
        t, s, w = np.empty((rows, 4), np.float32), np.empty((rows, 4), np.float32), np.empty(rows, np.uint8)
        k, gen = C.c_int(), C.c_int64()
        check(lib().rb_dataset_save_host(self._h, ob.ctypes.data, rew.ctypes.data, t.ctypes.data, s.ctypes.data, w.ctypes.data, C.byref(k), C.byref(gen)))
        return dict(ob=torch.from_numpy(ob), rew=torch.from_numpy(rew), t=torch.from_numpy(t), s=torch.from_numpy(s), stepped_with=torch.from_numpy(w),
                    step=k.value, generations=gen.value, draw=self._draw, dumped=self._dumped, num_envs=self.num_envs, rows=rows)

    def load_state_dict(self, sd):
        rows = int(lib().rb_dataset_ring_rows(self._h))
        if int(sd["rows"]) != rows or int(sd["num_envs"]) != self.num_envs:
            raise ValueError("Dataset checkpoint is for %d envs / %d ring rows" % (sd["num_envs"], sd["rows"]))
        a = [np.ascontiguousarray(sd[k].numpy()) for k in ("ob", "rew", "t", "s", "stepped_with")]
        check(lib().rb_dataset_load_host(self._h, *[x.ctypes.data for x in a], int(sd["step"]), int(sd["generations"])))
        self._draw, self._dumped = int(sd["draw"]), int(sd["dumped"])

    def close(self):
        if getattr(self, "_h", None):
            lib().rb_dataset_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
