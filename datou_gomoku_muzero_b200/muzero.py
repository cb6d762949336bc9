"""Device-resident MuZero-mode search (reference MuZeroMCTS.search, mcts.py:288-362) for G games.

The tree lives in the engine (gmz_select_mz / gmz_expand_backup); hidden states live in a pool on the
device, one row per tree node, addressed by the slots the select kernel emits -- the reference instead
pickles every hidden state through two mp.Queue hops per batch (mcts.py:77-85, workers.py:357-369).
One loop iteration = one of the reference's batches: the len(survivors) identical selections are
deduplicated to one evaluation, the kernel applies that many backups.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from .network import GomokuNetEZ, _fold, _support_scalar

_M64 = (1 << 64) - 1


def evals_per_search(num_simulations, num_top_actions, n_survivors):
    """Number of recurrent evaluations one search performs when the initial survivor list has
    `n_survivors` entries: the halving state machine of mcts.py:158-181 driven with sim_count += k."""
    S, K = int(num_simulations), int(num_top_actions)
    lg2 = np.log2(K) if K > 1 else 0.0
    nxt = S if (K <= 1 or lg2 <= 0) else int(min(np.floor(S / (lg2 * K)) * K, S))
    m, used, sims, k, evals = K, 0.0, 1, int(n_survivors), 0
    while sims < S and k > 0:
        sims += k
        evals += 1
        if sims >= nxt:
            m //= 2
            if m >= 1:
                extra = (S - used) if (m <= 1 or lg2 <= 0) else np.floor(S / (lg2 * m)) * m
                used += extra
                nxt = min(nxt + int(extra), S)
                k = min(m, k)
    return evals


class MuZeroDeviceSearch:
    """initial_fn(obs f32 [G,3,N,N]) -> (logits f32 [G,A], values [G], hidden [G, ...])
    recurrent_fn(hidden [G, ...], actions int64 [G]) -> (logits, values, rewards [G], next_hidden [G, ...])
    Hidden rows may be any dtype / shape; the pool is allocated from the first initial_fn result."""

    def __init__(self, engine, initial_fn, recurrent_fn, nodes_per_game=None):
        if engine.mode != "MuZero":
            raise ValueError("MuZeroDeviceSearch needs an engine created with mode='MuZero'")
        self.e, self.initial_fn, self.recurrent_fn = engine, initial_fn, recurrent_fn
        self.nodes = int(nodes_per_game) if nodes_per_game else engine.S
        self.pool = None
        self.evaluations = 0

    def _ensure_pool(self, hidden):
        shape = (self.e.G * self.nodes + 1,) + tuple(hidden.shape[1:])      # +1: dummy row for idle games
        if self.pool is None or self.pool.shape != shape or self.pool.dtype != hidden.dtype:
            if hidden.dim() == 4 and hidden.is_contiguous(memory_format=torch.channels_last):
                self.pool = torch.empty(shape, dtype=hidden.dtype, device=hidden.device).contiguous(
                    memory_format=torch.channels_last)
            else:
                self.pool = torch.empty(shape, dtype=hidden.dtype, device=hidden.device)
        return self.pool

    def _row(self, slot):
        """engine slot (g*S + node) -> pool row (g*nodes + node); -1 -> the dummy row."""
        e = self.e
        s = slot.long()
        row = (s // e.S) * self.nodes + (s % e.S)
        return torch.where(s < 0, torch.full_like(row, e.G * self.nodes), row)

    @torch.no_grad()
    def search(self, gumbel, max_steps=None):
        """Runs one search for every game from the engine's current roots; call engine.finalize() after."""
        e = self.e
        lg, v, h = self.initial_fn(e.root_obs())
        pool = self._ensure_pool(h)
        base = torch.arange(e.G, device=e.device) * self.nodes
        pool.index_copy_(0, base, h)
        e.root_expand(lg, v, gumbel)
        steps, limit = 0, int(max_steps) if max_steps else e.S - 1      # at most S-1 evaluations (sim_count starts at 1)
        while steps < limit:
            parent, action, child, depth = e.select_mz()
            if steps % 8 == 7 and int(parent.max().item()) < 0:       # every game is done
                break
            if steps + 2 > self.nodes:          # this step may create node id steps + 1
                raise RuntimeError("hidden-state pool too small: pass a larger nodes_per_game")
            h_in = pool.index_select(0, self._row(parent))
            lg, v, r, h_out = self.recurrent_fn(h_in, action.long().clamp_min(0))
            pool.index_copy_(0, self._row(child), h_out)
            e.expand_backup(lg, v, r)
            steps += 1
            self.evaluations += 1
        return steps


class TorchE0:
    """The fixed evaluator E0 in MuZero mode, as torch integer ops on the device (hidden state = the
    64-bit hash, one int64 per node).  Same integers as the CUDA E0 (csrc/gmz_common.cuh) and tests/golden/e0_py.py."""

    GOLD, CV, CA, CR = 0x9E3779B97F4A7C15, 0xD1B54A32D192ED03, 0x8CB92BA72F3D8DD7, 0xA24BAED4963EE407

    def __init__(self, board_size, seed=0, logit_div=16, device="cuda"):
        self.N, self.A = board_size, board_size * board_size
        self.seed, self.div, self.device = int(seed), float(logit_div), device
        self.a1 = self._c(((np.arange(1, self.A + 1, dtype=np.uint64) * np.uint64(self.GOLD))).astype(np.int64))

    def _c(self, x):
        return torch.as_tensor(x, dtype=torch.int64, device=self.device)

    @staticmethod
    def _s(v):          # python int (unsigned 64) -> signed int64 value
        v &= _M64
        return v - (1 << 64) if v >= (1 << 63) else v

    @staticmethod
    def _shr(z, k):     # logical right shift on int64
        return (z >> k) & ((1 << (64 - k)) - 1)

    def mix(self, z):
        z = z ^ self._shr(z, 30)
        z = z * self._s(0xBF58476D1CE4E5B9)
        z = z ^ self._shr(z, 27)
        z = z * self._s(0x94D049BB133111EB)
        return z ^ self._shr(z, 31)

    def heads(self, h):
        k = self._shr(self.mix(h[:, None] + self.a1[None, :]), 58)
        logits = (k - 32).to(torch.float32) / self.div
        value = ((self._shr(self.mix(h ^ self._s(self.CV)), 40) % 33) - 16).to(torch.float64) / 16.0
        return logits, value

    def initial(self, obs):
        B, A = obs.shape[0], self.A
        nw = (A + 63) // 64
        planes = (obs.reshape(B, 3, A) > 0.5)
        weights = torch.zeros((nw, A), dtype=torch.int64, device=obs.device)
        idx = torch.arange(A, device=obs.device)
        bit = torch.ones(A, dtype=torch.int64, device=obs.device) << (idx % 64)      # bit 63 wraps to the sign bit
        weights[idx // 64, idx] = bit
        own = (planes[:, 0, None, :].long() * weights[None]).sum(-1)                # disjoint bits: sum == or
        opp = (planes[:, 1, None, :].long() * weights[None]).sum(-1)
        has_last = planes[:, 2].any(dim=1)
        last = torch.where(has_last, planes[:, 2].float().argmax(dim=1), torch.full((B,), -1, device=obs.device))
        h = torch.full((B,), self._s(self.mix_int(self.seed ^ self.GOLD)), dtype=torch.int64, device=obs.device)
        for w in range(nw):
            h = self.mix(h ^ own[:, w])
        for w in range(nw):
            h = self.mix(h ^ opp[:, w])
        h = self.mix(h ^ (last + 1))
        lg, v = self.heads(h)
        return lg, v, h

    @staticmethod
    def mix_int(z):
        z &= _M64
        z ^= z >> 30; z = (z * 0xBF58476D1CE4E5B9) & _M64
        z ^= z >> 27; z = (z * 0x94D049BB133111EB) & _M64
        return z ^ (z >> 31)

    def recurrent(self, h_parent, actions):
        hc = self.mix(h_parent ^ self.mix(actions + 1 + self._s(self.CA)))
        lg, v = self.heads(hc)
        r = ((self._shr(self.mix(hc ^ self._s(self.CR)), 40) % 5) - 2).to(torch.float64) / 16.0
        return lg, v, r, hc


class FoldedRecurrentInference:
    """`recurrent_inference` (network.py:145-152) for the engine: BatchNorm folded, cuDNN fused
    conv+bias(+residual)+ReLU, channels_last bf16 hidden states, reward head weights permuted to the
    NHWC flatten order so the [B, C*N*N] view of the hidden state is free.  Library kernels."""

    def __init__(self, net: GomokuNetEZ, dtype=torch.bfloat16):
        d, p = net.dynamics_net, net.prediction_net
        self.dtype = dtype
        self.embed = d.action_embed_conv.weight.detach().to(dtype).reshape(-1)           # [16]
        self.stem = _fold(d.conv, d.bn, dtype)
        self.blocks = [(_fold(b.conv1, b.bn1, dtype), _fold(b.conv2, b.bn2, dtype)) for b in d.resblocks]
        n, C = net.board_size, d.conv.out_channels
        w1 = d.reward_fc[0].weight.detach()                                                # [hid, C*n*n] in CHW order
        self.r1 = (w1.reshape(-1, C, n, n).permute(0, 2, 3, 1).reshape(w1.shape[0], -1).to(dtype).contiguous(),
                   d.reward_fc[0].bias.detach().to(dtype))
        self.r2 = (d.reward_fc[2].weight.detach().to(dtype), d.reward_fc[2].bias.detach().to(dtype))
        self.pol, self.val = _fold(p.policy_conv, p.policy_bn, dtype), _fold(p.value_conv, p.value_bn, dtype)
        self.policy_fc = (p.policy_fc.weight.detach().to(dtype), p.policy_fc.bias.detach().to(dtype))
        self.value_fc1 = (p.value_fc1.weight.detach().to(dtype), p.value_fc1.bias.detach().to(dtype))
        self.value_fc2 = (p.value_fc2.weight.detach().to(dtype), p.value_fc2.bias.detach().to(dtype))
        self.v_sup, self.r_sup = net.v_sup, net.r_sup
        self.n = n

    @staticmethod
    def _cr(x, wb, pad):
        return torch.cudnn_convolution_relu(x, wb[0], wb[1], (1, 1), (pad, pad), (1, 1), 1)

    @torch.no_grad()
    def __call__(self, hidden, actions):
        B, n = hidden.shape[0], self.n
        plane = F.one_hot(actions, n * n).to(self.dtype).reshape(B, 1, n, n)
        emb = plane * self.embed.reshape(1, -1, 1, 1)                                      # 1x1 conv of a one-hot plane
        x = torch.cat((hidden, emb), dim=1).contiguous(memory_format=torch.channels_last)
        h = self._cr(x, self.stem, 1)
        for c1, c2 in self.blocks:
            h = torch.cudnn_convolution_add_relu(self._cr(h, c1, 1), c2[0], h, 1.0, c2[1], (1, 1), (1, 1), (1, 1), 1)
        flat = h.permute(0, 2, 3, 1).reshape(B, -1)
        rew = _support_scalar(F.linear(F.relu(F.linear(flat, *self.r1)), *self.r2).float(), *self.r_sup)
        pl = self._cr(h, self.pol, 0).reshape(B, -1)
        vl = self._cr(h, self.val, 0).reshape(B, -1)
        logits = F.linear(pl, *self.policy_fc).float()
        value = _support_scalar(F.linear(F.relu(F.linear(vl, *self.value_fc1)), *self.value_fc2).float(), *self.v_sup)
        return logits.contiguous(), value.reshape(-1), rew.reshape(-1), h
