"""Dataset -- the reference's rollout buffer (/root/reference src/distilation/dataset.py:72-296) on the device.

Same method names: write / flush / training_batches / test_batch / num_episodes / last_step.  N lock-step envs append one
record each per call; sampling uses Philox (seed, draw counter) instead of Python's `random`.  The gzip-JSON page store
(DatasetStore, dataset.py:14-65: dump / pages / switch) is the next row of the build and raises here.
"""
import ctypes as C

import torch

from ._lib import check, lib, ptr, stream_ptr
from .config import EPISODE_STEPS, LSTM_BATCH_SIZE, STEPS_UNROLLED, TRAINING_EPOCHS

__all__ = ["Dataset"]


class Dataset:
    def __init__(self, dir_path=None, num_envs=1, generations=64, device=0, seed=0):
        self.dir_path, self.num_envs, self.seed = dir_path, int(num_envs), int(seed)
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        h = C.c_void_p()
        check(lib().rb_dataset_create(C.byref(h), self.num_envs, int(generations), self.device.index))
        self._h = h
        self._draw = 0

    # ---- recording -----------------------------------------------------------------------------------------
    def _dev(self, a, width):
        if a is None:
            return None
        t = torch.as_tensor(a, dtype=torch.float32).to(self.device).reshape(self.num_envs, width) if width else \
            torch.as_tensor(a, dtype=torch.float32).to(self.device).reshape(self.num_envs)
        return t.contiguous()

    def write(self, ob, reward=None, t_pdflat=None, s_pdflat=None, stepped_with="t"):
        """dataset.py:118-143.  ob [N,11], reward [N], pdflats [N,4] (device tensors or array-likes)."""
        ob, rw, t, s = self._dev(ob, 11), self._dev(reward, 0), self._dev(t_pdflat, 4), self._dev(s_pdflat, 4)
        check(lib().rb_dataset_write(self._h, ptr(ob), ptr(rw), ptr(t), ptr(s), 0 if stepped_with == "t" else 1, stream_ptr()))
        if any(x is not None and not torch.is_tensor(a) for x, a in ((ob, ob), (rw, reward), (t, t_pdflat), (s, s_pdflat))):
            torch.cuda.current_stream().synchronize()        # temporaries made from host data must outlive the copy

    def flush(self):
        check(lib().rb_dataset_flush(self._h))

    def num_episodes(self):
        return int(lib().rb_dataset_num_episodes(self._h))

    def last_step(self):
        return int(lib().rb_dataset_episode_len(self._h)) - 1

    # ---- batches -------------------------------------------------------------------------------------------
    def training_batch(self, batch_size=LSTM_BATCH_SIZE, steps=STEPS_UNROLLED, draw=None, with_indices=False):
        """One batch of dataset.py:184-194: (ob[T,B,11], t[T,B,4], prev[T,B,4], prew[T,B,1]) as device tensors."""
        if draw is None:
            draw, self._draw = self._draw, self._draw + 1
        dev = self.device
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

    # ---- page store: next row of the build -----------------------------------------------------------------
    def dump(self):
        raise NotImplementedError("gzip-JSON page store (dataset.py:14-65) is the next row of the build; the buffer is device-resident")

    pages = switch = dump

    def close(self):
        if getattr(self, "_h", None):
            lib().rb_dataset_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
