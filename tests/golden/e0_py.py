"""E0, the fixed deterministic evaluator (DESIGN.md "E0"), in pure Python/NumPy, speaking the
reference's inference-queue protocol (mcts.py:73-85, tests/test_mcts_logic.py:26-58).

Used (a) by make_golden.py to drive the imported reference and (b) by CPU tests to check
that the C oracle's and the CUDA engine's E0 produce the same integers.  Values are returned
as Python floats / float64 so the reference accumulates in float64.
"""
from __future__ import annotations

from queue import Empty

import numpy as np

M64 = (1 << 64) - 1
GOLD = 0x9E3779B97F4A7C15
CV = 0xD1B54A32D192ED03
CA = 0x8CB92BA72F3D8DD7
CR = 0xA24BAED4963EE407


def mix64(z: int) -> int:
    z &= M64
    z ^= z >> 30
    z = (z * 0xBF58476D1CE4E5B9) & M64
    z ^= z >> 27
    z = (z * 0x94D049BB133111EB) & M64
    z ^= z >> 31
    return z


def _mix64_vec(z: np.ndarray) -> np.ndarray:
    z = z.astype(np.uint64)
    with np.errstate(over="ignore"):
        z ^= z >> np.uint64(30)
        z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27)
        z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    return z


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
    h = mix64((seed & M64) ^ GOLD)
    for w in own:
        h = mix64(h ^ w)
    for w in opp:
        h = mix64(h ^ w)
    return mix64(h ^ ((last + 1) & M64))


def heads(h: int, A: int, logit_div: int):
    a = np.arange(1, A + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(h) + a * np.uint64(GOLD)
    k = (_mix64_vec(z) >> np.uint64(58)).astype(np.int64)
    logits = (k - 32).astype(np.float32) / np.float32(logit_div)
    value = float(((mix64(h ^ CV) >> 40) % 33) - 16) / 16.0
    return logits.astype(np.float32), value


def child_hidden(h_parent: int, action: int) -> int:
    return mix64(h_parent ^ mix64(((action + 1) + CA) & M64))


def reward_of(h: int) -> float:
    return float(((mix64(h ^ CR) >> 40) % 5) - 2) / 16.0


class E0Queue:
    """Synchronous stand-in for (request_queue, result_queue), same shape as the reference's
    MockInferenceQueue / LocalInferenceEngine.  kind 0 = hash evaluator, 1 = constant."""

    def __init__(self, seed=0, logit_div=16, kind=0, const_value=0.5, const_reward=0.0):
        self.seed, self.logit_div, self.kind = int(seed), int(logit_div), int(kind)
        self.const_value, self.const_reward = float(const_value), float(const_reward)
        self.pending = []
        self.n_initial = 0
        self.n_recurrent = 0

    def put(self, item):
        self.pending.append(item)

    def get_nowait(self):
        if not self.pending:
            raise Empty
        return self.pending.pop(0)

    def get(self, timeout=None):
        if not self.pending:
            raise Empty
        _, kind, data = self.pending.pop(0)
        if kind == "initial":
            self.n_initial += 1
            A = data.shape[1] * data.shape[2]
            if self.kind == 1:
                return np.zeros(A, np.float32), self.const_value, np.array([[1]], np.uint64)
            h = hash_obs(data, self.seed)
            logits, value = heads(h, A, self.logit_div)
            return logits, value, np.array([[h]], dtype=np.uint64)
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
            ps.append(logits); vs.append(value); hs.append(hc); rs.append(reward_of(hc))
        return (np.stack(ps), np.array(vs, np.float64).reshape(k, 1),
                np.array(hs, np.uint64).reshape(k, 1), np.array(rs, np.float64).reshape(k, 1))

    _A = 0

    def set_action_space(self, A):
        self._A = int(A)
