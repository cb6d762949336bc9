"""Prioritized replay with the SumTree resident on the GPU, API-compatible with the reference's
replay_buffer.py (SumTree :4-41, InMemoryReplayBuffer :43-106).  Sampling and priority updates are
CUDA kernels (csrc/gmz_per.cu) that keep the reference's sequential float64 semantics bit for bit.

Two buffers:
  * `InMemoryReplayBuffer` -- the reference's class: slices are host objects in `self.data` exactly as in
    the reference, only the tree lives on the device.
  * `DeviceReplayBuffer` -- the same ring + tree with the DATA on the device too: a ring of packed move
    records (csrc/gmz_records.cu) written by the self-play engine's trajectory hand-off.  `sample()` returns
    the trainer's batch tuple (workers.py:430-433) as device tensors; nothing touches the host between a
    finished game and a training batch.
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

    def _dev(self, x, dtype):
        """numpy / list / tensor -> contiguous device tensor of `dtype` (no copy if it already is one)."""
        if not torch.is_tensor(x):
            x = torch.as_tensor(np.ascontiguousarray(x))
        return x.to(device=self.device, dtype=dtype).contiguous()

    # ---- device-tensor entry points (no host round trip)
    def update_many(self, tree_idx, priorities):
        """Apply tree.update(idx_i, p_i) for i = 0..n-1 in order (replay_buffer.py:16-19)."""
        idx, pr = self._dev(tree_idx, torch.int64), self._dev(priorities, torch.float64)
        check(self.lib.gmz_per_update(_ptr(self.tree), self.capacity, _ptr(idx), _ptr(pr), int(idx.numel()),
                                      self._stream()), "gmz_per_update")

    def update(self, tree_idx, priority):
        self.update_many([int(tree_idx)], [float(priority)])

    def add_many(self, priorities):
        pr = self._dev(priorities, torch.float64)
        n = int(pr.numel())
        if n == 0:
            return
        scratch = torch.empty(n, dtype=torch.int64, device=self.device)
        check(self.lib.gmz_per_add(_ptr(self.tree), self.capacity, self.write_ptr, _ptr(pr), n, _ptr(scratch),
                                   self._stream()), "gmz_per_add")
        self.write_ptr = (self.write_ptr + n) % self.capacity
        self.count = min(self.capacity, self.count + n)

    def add(self, priority):
        self.add_many([float(priority)])

    def sample_device(self, u01, beta):
        """Stratified draw on the device: (tree_idx int64 [B], priority f64 [B], is_weights f32 [B]) device tensors."""
        u = self._dev(u01, torch.float64)
        B = int(u.numel())
        idx = torch.empty(B, dtype=torch.int64, device=self.device)
        pr = torch.empty(B, dtype=torch.float64, device=self.device)
        w = torch.empty(B, dtype=torch.float32, device=self.device)
        check(self.lib.gmz_per_sample(_ptr(self.tree), self.capacity, self.count, _ptr(u), B, float(beta),
                                      _ptr(idx), _ptr(pr), _ptr(w), self._stream()), "gmz_per_sample")
        return idx, pr, w

    def sample(self, u01, beta):
        """Host-facing form of sample_device (numpy results)."""
        idx, pr, w = self.sample_device(u01, beta)
        return idx.cpu().numpy(), pr.cpu().numpy(), w.cpu().numpy()

    def get_leaf(self, value):
        """SumTree.get_leaf (replay_buffer.py:27-38): the descent runs as tensor ops on the device, one
        synchronisation at the end (not one per level)."""
        tree, n = self.tree, self.tree.numel()
        parent = torch.zeros((), dtype=torch.int64, device=self.device)
        v = torch.tensor(float(value), dtype=torch.float64, device=self.device)
        levels = 0
        while (1 << levels) - 1 < n:        # deepest possible leaf level
            levels += 1
        for _ in range(levels):
            left = 2 * parent + 1
            inner = left < n
            lv = tree[left.clamp(max=n - 1)]
            go_left = v <= lv
            nxt = torch.where(go_left, left, left + 1)
            v = torch.where(inner & ~go_left, v - lv, v)
            parent = torch.where(inner, nxt, parent)
        return int(parent.item())

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


class DeviceReplayBuffer:
    """The reference's InMemoryReplayBuffer (replay_buffer.py:43-106) with tree AND data on the device.

    data ring : `capacity` packed move records (include/gmz.h gmz_move_record); record i is the TrainingSlice the
                reference would have stored at data[i] -- a slice is that record plus the next U of the same game, which
                follow it in the ring (games are appended whole, in move order; a ring overwrites oldest first, so the
                successors of a live record are live).
    add_packed(games)  = `for s in slices: buffer.add(s)` for every game of a trajectory.PackedGames (workers.py:399-407)
    sample(B)          -> ((obs, act, rew, pi, val), tree_indices, is_weights), all device tensors: the tuple
                          data_loader_worker stacks (workers.py:430-433); rot_k / flip apply the trainer's D4
                          augmentation while gathering (loss.py:37-51)
    update_priorities  = replay_buffer.py:98-103 on device tensors (td_errors float32, as the loss returns them)
    rewrite_targets    = the re-analysis write-back (db_manager.py:189-214) for records in the ring
    """

    def __init__(self, capacity, board_size, device=None, unroll_steps=None):
        self.sum_tree = SumTree(capacity, device)
        self.capacity, self.N, self.A = int(capacity), int(board_size), int(board_size) ** 2
        self.device, self.lib = self.sum_tree.device, self.sum_tree.lib
        self.stride = int(self.lib.gmz_move_record_bytes(self.N))
        self.ring = torch.zeros((self.capacity, self.stride), dtype=torch.uint8, device=self.device)
        self.max_priority = torch.ones((), dtype=torch.float64, device=self.device)
        self.unroll = int(config.NUM_UNROLL_STEPS if unroll_steps is None else unroll_steps)

    def add_packed(self, games):
        rec = games.records
        if rec.device != self.device:
            rec = rec.to(self.device)
        total = int(rec.shape[0])
        for s0 in range(0, total, self.capacity):        # more records than the ring holds: wrap exactly like sequential add()s
            part = rec[s0:s0 + self.capacity]
            M = int(part.shape[0])
            wp = self.sum_tree.write_ptr
            first = min(M, self.capacity - wp)
            self.ring[wp:wp + first].copy_(part[:first])
            if first < M:
                self.ring[:M - first].copy_(part[first:])
            pr = (self.max_priority if config.ENABLE_PER else torch.ones_like(self.max_priority)).expand(M)
            self.sum_tree.add_many(pr)

    def sample(self, batch_size, u01=None, rot_k=0, flip=False):
        if self.sum_tree.count < batch_size:
            return None, None, None
        B = int(batch_size)
        if config.ENABLE_PER:
            u = torch.rand(B, dtype=torch.float64, device=self.device) if u01 is None else u01
            idx, _, w = self.sum_tree.sample_device(u, config.PER_BETA)
            pos = idx - (self.capacity - 1)
        else:
            pos = torch.randperm(self.sum_tree.count, device=self.device)[:B]
            idx, w = pos, torch.ones(B, dtype=torch.float32, device=self.device)
        return self.batch(pos, rot_k, flip), idx, w

    def batch(self, positions, rot_k=0, flip=False):
        """Trainer tuple for the slices stored at ring `positions` (int64 device tensor)."""
        pos = self.sum_tree._dev(positions, torch.int64)
        B, U, N, A, dev = int(pos.numel()), self.unroll, self.N, self.A, self.device
        obs = torch.empty((B, U + 1, 3, N, N), dtype=torch.float32, device=dev)
        act = torch.empty((B, U), dtype=torch.int32, device=dev)
        rew = torch.empty((B, U), dtype=torch.float32, device=dev)
        pi = torch.empty((B, U + 1, A), dtype=torch.float64, device=dev)
        val = torch.empty((B, U + 1), dtype=torch.float32, device=dev)
        check(self.lib.gmz_records_batch(_ptr(self.ring), self.capacity, N, _ptr(pos), B, U, int(rot_k) % 4, 1 if flip else 0,
                                         _ptr(obs), _ptr(act), _ptr(rew), _ptr(pi), _ptr(val), self.sum_tree._stream()),
              "gmz_records_batch")
        return obs, act, rew, pi, val

    def rewrite_targets(self, positions, policies, value_targets):
        """Re-analysis write-back (db_manager.py:189-214) for records resident in the ring: the policy and the value
        target of the move stored at each ring position are replaced; every slice that covers the move -- slices are
        assembled from consecutive records when a batch is built -- then carries the new window.  positions int64 [M],
        policies float64 [M, A], value_targets float32 [M] (host or device)."""
        pos = self.sum_tree._dev(positions, torch.int64).reshape(-1)
        pol = self.sum_tree._dev(policies, torch.float64).reshape(pos.numel(), self.A).contiguous()
        val = self.sum_tree._dev(value_targets, torch.float32).reshape(pos.numel()).contiguous()
        if pos.numel() == 0:
            return
        if int(pos.min()) < 0 or int(pos.max()) >= self.capacity:
            raise ValueError("ring position out of range")
        self.ring[pos, 64:64 + 8 * self.A] = pol.view(torch.uint8).reshape(pos.numel(), 8 * self.A)      # gmz_move_record: policy f64[A]
        self.ring[pos, 36:40] = val.view(torch.uint8).reshape(pos.numel(), 4)                             # header: value_target f32

    def update_priorities(self, tree_indices, td_errors):
        if not config.ENABLE_PER:
            return
        td = self.sum_tree._dev(td_errors, torch.float32)
        if td.numel() == 0:
            return
        pr = (td.abs() + np.float32(config.PER_EPSILON)).double()        # float32 arithmetic like the reference, then widened
        self.max_priority = torch.maximum(self.max_priority, pr.max())
        self.sum_tree.update_many(tree_indices, pr)

    def __len__(self):
        return self.sum_tree.count
