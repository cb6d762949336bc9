"""E0, the fixed deterministic evaluator (DESIGN.md "E0"), in pure Python/NumPy, speaking the
reference's inference-queue protocol (mcts.py:73-85, tests/test_mcts_logic.py:26-58).

Used (a) by make_golden.py to drive the imported reference and (b) by CPU tests to check
that the C oracle's and the CUDA engine's E0 produce the same integers.

Definition (round 2, third version; cheap on a GPU: the board hash is order-free over the words with ONE
odd-constant multiply per plane word -- a bijection of the word, so two positions differing in one word never
collide before the final mix, and the final mix64 gives the avalanche -- and the per-action hash is 32-bit with
one multiply, its top bits being what the heads use):
    h0 = mix64(seed ^ GOLD)
    h  = mix64( XOR_w [ ((own_w ^ h0) + (2w+1) GOLD) K1  ^  rot32(((opp_w ^ h0) + (2w+2) GOLD) K2) ]  +  (last+1) CV )
         K1, K2 = mix64's two multipliers, rot32 = swap of the 32-bit halves
    s  = lo32(h) ^ hi32(h);   y = s + (a+1) * 0x9E3779B1;   x_a = (y ^ (y >> 16)) * 0x7FEB352D   (mod 2^32)
    logit_div > 0 (quantised):  logit_a = ((x_a >> 26) - 32) / logit_div       value = (((h >> 40) % 33) - 16) / 16
                                reward  = ((((h >> 16) & 0xFFFFFF) % 5) - 2) / 16
    logit_div = 0 (dense):      logit_a = ((x_a >> 8) - 2^23) * 2^-21  in [-4, 4)   -- 24 random mantissa bits
                                value   = ((h >> 40) - 2^23) * 2^-23   in [-1, 1)
                                reward  = (((h >> 16) & 0xFFFFFF) - 2^23) * 2^-25 in [-0.25, 0.25)
    MuZero mode: h_child = mix64(h_parent + (a+1) CA)
Every value / reward is exactly representable in float32, so the evaluator can hand the search
Python floats (the upstream test mock, float64 accumulation) or np.float32 scalars (what the
reference's inference server returns, workers.py:355,368 -- float32 accumulation under NumPy >= 2,
SURVEY.md App. A.7) without changing the numbers themselves: `value_dtype`.
"""
from __future__ import annotations

from queue import Empty

import numpy as np

M64 = (1 << 64) - 1
GOLD = 0x9E3779B97F4A7C15
CV = 0xD1B54A32D192ED03
CA = 0x8CB92BA72F3D8DD7
K1, K2 = 0xBF58476D1CE4E5B9, 0x94D049BB133111EB
GOLD32, M1_32 = 0x9E3779B1, 0x7FEB352D


def mix64(z: int) -> int:
    z &= M64
    z ^= z >> 30
    z = (z * K1) & M64
    z ^= z >> 27
    z = (z * K2) & M64
    z ^= z >> 31
    return z


def rot32(z: int) -> int:
    return ((z << 32) | (z >> 32)) & M64


def _words(plane_bits: np.ndarray, nw: int):
    x = int.from_bytes(np.packbits(plane_bits.astype(np.uint8), bitorder="little").tobytes(), "little")
    return [(x >> (64 * w)) & M64 for w in range(nw)]


def hash_obs(obs: np.ndarray, seed: int) -> int:
    """obs = float32 [3,N,N] planes own / opp / last-move one-hot (game.py:12-17)."""
    A = obs.shape[1] * obs.shape[2]
    nw = (A + 63) // 64
    own = _words(obs[0].reshape(-1) > 0.5, nw)
    opp = _words(obs[1].reshape(-1) > 0.5, nw)
    lm = np.flatnonzero(obs[2].reshape(-1) > 0.5)
    last = int(lm[0]) if len(lm) else -1
    h0 = mix64((seed & M64) ^ GOLD)
    acc = 0
    for w in range(nw):
        acc ^= (((own[w] ^ h0) + (2 * w + 1) * GOLD) * K1) & M64
        acc ^= rot32((((opp[w] ^ h0) + (2 * w + 2) * GOLD) * K2) & M64)
    return mix64(acc + (last + 1) * CV)


def action_hash(h: int, A: int) -> np.ndarray:
    """x_a for a = 0..A-1 (uint32)."""
    s = np.uint32((h & 0xFFFFFFFF) ^ (h >> 32))
    with np.errstate(over="ignore"):
        x = s + np.arange(1, A + 1, dtype=np.uint32) * np.uint32(GOLD32)
        x ^= x >> np.uint32(16)
        x *= np.uint32(M1_32)
    return x


def heads(h: int, A: int, logit_div: int):
    """(logits float32 [A], value as a Python float)."""
    x = action_hash(h, A)
    vk = (h >> 40) & 0xFFFFFF
    if logit_div > 0:
        logits = ((x >> np.uint32(26)).astype(np.int64) - 32).astype(np.float32) / np.float32(logit_div)
        value = float((vk % 33) - 16) / 16.0
    else:
        logits = ((x >> np.uint32(8)).astype(np.int64) - (1 << 23)).astype(np.float32) * np.float32(2.0 ** -21)
        value = float(vk - (1 << 23)) * 2.0 ** -23
    return logits.astype(np.float32), value


def child_hidden(h_parent: int, action: int) -> int:
    return mix64(h_parent + (action + 1) * CA)


def reward_of(h: int, logit_div: int = 16) -> float:
    rk = (h >> 16) & 0xFFFFFF
    if logit_div > 0:
        return float((rk % 5) - 2) / 16.0
    return float(rk - (1 << 23)) * 2.0 ** -25


class E0Queue:
    """Synchronous stand-in for (request_queue, result_queue), same shape as the reference's
    MockInferenceQueue / LocalInferenceEngine.  kind 0 = hash evaluator (logit_div 0 = dense), 1 = constant.
    value_dtype: None -> Python floats / float64 arrays (the upstream test mock); np.float32 -> np.float32
    scalars and float32 [k,1] arrays (the reference's inference server, workers.py:355,368)."""

    def __init__(self, seed=0, logit_div=16, kind=0, const_value=0.5, const_reward=0.0, value_dtype=None):
        self.seed, self.logit_div, self.kind = int(seed), int(logit_div), int(kind)
        self.const_value, self.const_reward = float(const_value), float(const_reward)
        self.f32 = value_dtype is not None and np.dtype(value_dtype) == np.float32
        self.pending = []
        self.n_initial = 0
        self.n_recurrent = 0

    def put(self, item):
        self.pending.append(item)

    def get_nowait(self):
        if not self.pending:
            raise Empty
        return self.pending.pop(0)

    def _scalar(self, v):
        return np.float32(v) if self.f32 else float(v)

    def get(self, timeout=None):
        if not self.pending:
            raise Empty
        _, kind, data = self.pending.pop(0)
        if kind == "initial":
            self.n_initial += 1
            A = data.shape[1] * data.shape[2]
            if self.kind == 1:
                return np.zeros(A, np.float32), self._scalar(self.const_value), np.array([[1]], np.uint64)
            h = hash_obs(data, self.seed)
            logits, value = heads(h, A, self.logit_div)
            return logits, self._scalar(value), np.array([[h]], dtype=np.uint64)
        hidden, actions = data
        self.n_recurrent += 1
        k = len(actions)
        ps, vs, hs, rs = [], [], [], []
        for i in range(k):
            if self.kind == 1:
                A = self._A
                ps.append(np.zeros(A, np.float32)); vs.append(self.const_value)
                hs.append(2); rs.append(self.const_reward)
                continue
            hc = child_hidden(int(hidden[i, 0]), int(actions[i]))
            logits, value = heads(hc, self._A, self.logit_div)
            ps.append(logits); vs.append(value); hs.append(hc); rs.append(reward_of(hc, self.logit_div))
        dt = np.float32 if self.f32 else np.float64
        return (np.stack(ps), np.array(vs, dt).reshape(k, 1),
                np.array(hs, np.uint64).reshape(k, 1), np.array(rs, dt).reshape(k, 1))

    _A = 0

    def set_action_space(self, A):
        self._A = int(A)
