"""Prioritized replay with the SumTree resident on the GPU, API-compatible with the reference's
replay_buffer.py (SumTree :4-41, InMemoryReplayBuffer :43-106).  Sampling and priority updates are
CUDA kernels (csrc/gmz_per.cu) that keep the reference's sequential float64 semantics bit for bit;
the slices themselves stay host objects exactly as in the reference (`self.data` list).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check
from .config import config


def _ptr(t):
    return C.c_void_p(t.data_ptr())


class SumTree:
    def __init__(self, capacity: int, device=None):
        if not torch.cuda.is_available():
            raise _lib.GmzError("SumTree needs a CUDA device (there is no CPU fallback)")
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.capacity = int(capacity)
        self.tree = torch.zeros(2 * self.capacity - 1, dtype=torch.float64, device=self.device)
        self.write_ptr = 0
        self.count = 0

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def update_many(self, tree_idx, priorities):
        """Apply tree.update(idx_i, p_i) for i = 0..n-1 in order (replay_buffer.py:16-19)."""
        idx = torch.as_tensor(np.ascontiguousarray(tree_idx, dtype=np.int64)).to(self.device)
        pr = torch.as_tensor(np.ascontiguousarray(priorities, dtype=np.float64)).to(self.device)
        check(self.lib.gmz_per_update(_ptr(self.tree), self.capacity, _ptr(idx), _ptr(pr), int(idx.numel()),
                                      self._stream()), "gmz_per_update")

    def update(self, tree_idx, priority):
        self.update_many([int(tree_idx)], [float(priority)])

    def add_many(self, priorities):
        n = len(priorities)
        pr = torch.as_tensor(np.ascontiguousarray(priorities, dtype=np.float64)).to(self.device)
        scratch = torch.empty(n, dtype=torch.int64, device=self.device)
        check(self.lib.gmz_per_add(_ptr(self.tree), self.capacity, self.write_ptr, _ptr(pr), n, _ptr(scratch),
                                   self._stream()), "gmz_per_add")
        self.write_ptr = (self.write_ptr + n) % self.capacity
        self.count = min(self.capacity, self.count + n)

    def add(self, priority):
        self.add_many([float(priority)])

    def sample(self, u01, beta):
        """Stratified draw: returns (tree_idx int64 [B], priority f64 [B], is_weights f32 [B]) on the host."""
        B = len(u01)
        u = torch.as_tensor(np.ascontiguousarray(u01, dtype=np.float64)).to(self.device)
        idx = torch.empty(B, dtype=torch.int64, device=self.device)
        pr = torch.empty(B, dtype=torch.float64, device=self.device)
        w = torch.empty(B, dtype=torch.float32, device=self.device)
        check(self.lib.gmz_per_sample(_ptr(self.tree), self.capacity, self.count, _ptr(u), B, float(beta),
                                      _ptr(idx), _ptr(pr), _ptr(w), self._stream()), "gmz_per_sample")
        return idx.cpu().numpy(), pr.cpu().numpy(), w.cpu().numpy()

    def get_leaf(self, value):
        # one-sample descent through the same kernel: segment = total, u = value / total is not
        # bit-safe, so walk on the host copy of the (tiny) path instead
        tree = self.tree
        parent, n = 0, tree.numel()
        value = float(value)
        while 2 * parent + 1 < n:
            left = 2 * parent + 1
            lv = float(tree[left])
            if value <= lv:
                parent = left
            else:
                value -= lv
                parent = left + 1
        return parent

    def total_priority(self):
        return float(self.tree[0])


class InMemoryReplayBuffer:
    def __init__(self, capacity, device=None):
        self.capacity = capacity
        self.sum_tree = SumTree(capacity, device)
        self.data = [None] * capacity
        self.max_priority = 1.0

    def add(self, training_slice):
        self.data[self.sum_tree.write_ptr] = training_slice
        self.sum_tree.add(self.max_priority if config.ENABLE_PER else 1.0)

    def add_many(self, slices):
        """Batch form of add(): one kernel for the whole game's slices (workers.py:399-407 adds them in a loop)."""
        p = self.max_priority if config.ENABLE_PER else 1.0
        wp = self.sum_tree.write_ptr
        for i, s in enumerate(slices):
            self.data[(wp + i) % self.capacity] = s
        self.sum_tree.add_many([p] * len(slices))

    def sample(self, batch_size):
        if self.sum_tree.count < batch_size:
            return None, None, None
        if config.ENABLE_PER:
            u = np.random.random_sample(batch_size)      # the doubles np.random.uniform would consume
            idx, _, w = self.sum_tree.sample(u, config.PER_BETA)
            batch = [self.data[int(i) - self.capacity + 1] for i in idx]
            return batch, [int(i) for i in idx], w
        indices = np.random.choice(self.sum_tree.count, batch_size, replace=False)
        return [self.data[i] for i in indices], indices, np.ones(batch_size, dtype=np.float32)

    def update_priorities(self, tree_indices, td_errors):
        if not config.ENABLE_PER:
            return
        priorities = np.abs(td_errors) + config.PER_EPSILON      # keeps td_errors' dtype (float32 from the loss)
        if len(priorities) == 0:
            return
        self.max_priority = max(self.max_priority, priorities.max())
        self.sum_tree.update_many(np.asarray(tree_indices, dtype=np.int64), priorities.astype(np.float64))

    def __len__(self):
        return self.sum_tree.count
