"""ctypes wrapper over oracle/libgmz_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  It is the checker, never the product: nothing in
datou_gomoku_muzero_b200/ imports it.  Parity of the C restatement is pinned
against tests/golden/*.npz (generated from the imported Python reference).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgmz_oracle.so")


class OrcConfig(C.Structure):
    _fields_ = [
        ("board_size", C.c_int32), ("n_in_row", C.c_int32),
        ("num_simulations", C.c_int32), ("num_top_actions", C.c_int32),
        ("mode", C.c_int32), ("eval_kind", C.c_int32), ("logit_div", C.c_int32), ("accum_dtype", C.c_int32),
        ("c_visit", C.c_double), ("c_scale", C.c_double), ("minmax_delta", C.c_double), ("discount", C.c_double),
        ("const_value", C.c_double), ("const_reward", C.c_double),
        ("eval_seed", C.c_uint64),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (seconds).  Building the checker is not using it."""
    src = os.path.join(_HERE, "gmz_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_sumtree_get_leaf.restype = C.c_int64
        _lib.orc_per_update.restype = C.c_double
    return _lib


def make_config(board_size=15, n_in_row=5, num_simulations=400, num_top_actions=16, mode=0,
                eval_kind=0, logit_div=16, c_visit=30.0, c_scale=1.0, minmax_delta=1e-3, discount=0.997,
                const_value=0.5, const_reward=0.0, eval_seed=0, accum_dtype=0) -> OrcConfig:
    """logit_div 0 = dense (unquantised) E0 logits / values; accum_dtype 1 = the evaluator returns np.float32
    scalars (the reference's inference server), i.e. float32 value_sum / Q / MinMaxStats."""
    return OrcConfig(board_size, n_in_row, num_simulations, num_top_actions, mode, eval_kind, logit_div, int(accum_dtype),
                     float(c_visit), float(c_scale), float(minmax_delta), float(discount),
                     float(const_value), float(const_reward), int(eval_seed) & 0xFFFFFFFFFFFFFFFF)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def search(cfg: OrcConfig, board, player, last_move, move_count, gumbel, trace=False):
    """One search.  last_move is an action index or -1 (None).  Returns a dict."""
    A = cfg.board_size ** 2
    S = cfg.num_simulations
    board = np.ascontiguousarray(board, dtype=np.int8).reshape(A)
    gumbel = np.ascontiguousarray(gumbel, dtype=np.float64).reshape(A)
    policy = np.zeros(A, np.float64)
    value = C.c_double(0.0)
    action = C.c_int32(-1)
    visits = np.zeros(A, np.int32)
    la = np.full(S + 1, -1, np.int32)
    ld = np.full(S + 1, -1, np.int32)
    counts = np.zeros(5, np.int32)
    mm = np.zeros(2, np.float64)
    rc = lib().orc_search(C.byref(cfg), _p(board, C.c_int8), int(player), int(last_move), int(move_count),
                          _p(gumbel, C.c_double), _p(policy, C.c_double), C.byref(value), C.byref(action),
                          _p(visits, C.c_int32), _p(la, C.c_int32), _p(ld, C.c_int32),
                          _p(counts, C.c_int32), _p(mm, C.c_double))
    out = dict(rc=rc, policy=policy, value=value.value, action=action.value, visits=visits,
               sim_count=int(counts[0]), n_evals=int(counts[1]), n_nodes=int(counts[2]), max_depth=int(counts[3]),
               all_visited=int(counts[4]),
               minmax=mm)
    if trace:
        out["leaf_actions"] = la[:counts[1]].copy()
        out["leaf_depths"] = ld[:counts[1]].copy()
    return out


def search_batch(cfg: OrcConfig, boards, players, last_moves, move_counts, gumbel, n_threads=0, want_visits=True):
    A = cfg.board_size ** 2
    boards = np.ascontiguousarray(boards, dtype=np.int8).reshape(-1, A)
    G = boards.shape[0]
    players = np.ascontiguousarray(players, dtype=np.int8).reshape(G)
    last_moves = np.ascontiguousarray(last_moves, dtype=np.int32).reshape(G)
    move_counts = np.ascontiguousarray(move_counts, dtype=np.int32).reshape(G)
    gumbel = np.ascontiguousarray(gumbel, dtype=np.float64).reshape(G, A)
    policy = np.zeros((G, A), np.float64)
    value = np.zeros(G, np.float64)
    action = np.zeros(G, np.int32)
    visits = np.zeros((G, A), np.int32) if want_visits else None
    lib().orc_search_batch(C.byref(cfg), G, _p(boards, C.c_int8), _p(players, C.c_int8), _p(last_moves, C.c_int32),
                           _p(move_counts, C.c_int32), _p(gumbel, C.c_double), _p(policy, C.c_double),
                           _p(value, C.c_double), _p(action, C.c_int32),
                           _p(visits, C.c_int32) if want_visits else None, int(n_threads))
    return policy, value, action, visits


_EVAL_FN = C.CFUNCTYPE(None, C.POINTER(C.c_int8), C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_double))
_eval_keepalive = None


def set_eval_callback(fn):
    """Evaluator for configs made with eval_kind=3 (AlphaZero mode, single-threaded `search` only):
    fn(board int8 [N,N], player, last_move) -> (logits float32 [A], value).  None removes it."""
    global _eval_keepalive
    if fn is None:
        lib().orc_set_eval_callback(_EVAL_FN(0))
        _eval_keepalive = None
        return

    def tramp(board_p, n, player, last_move, logits_p, value_p):
        A = n * n
        board = np.ctypeslib.as_array(board_p, shape=(A,)).reshape(n, n).copy()
        lg, v = fn(board, int(player), int(last_move))
        np.ctypeslib.as_array(logits_p, shape=(A,))[:] = np.asarray(lg, np.float32).reshape(A)
        value_p[0] = float(v)
    _eval_keepalive = _EVAL_FN(tramp)
    lib().orc_set_eval_callback(_eval_keepalive)


def max_threads() -> int:
    return int(lib().orc_max_threads())


def e0_initial(cfg: OrcConfig, board, player, last_move):
    A = cfg.board_size ** 2
    board = np.ascontiguousarray(board, dtype=np.int8).reshape(A)
    logits = np.zeros(A, np.float32)
    value = C.c_double()
    hidden = C.c_uint64()
    lib().orc_e0_initial(C.byref(cfg), _p(board, C.c_int8), int(player), int(last_move),
                         _p(logits, C.c_float), C.byref(value), C.byref(hidden))
    return logits, value.value, hidden.value


def e0_recurrent(cfg: OrcConfig, h_parent, action):
    A = cfg.board_size ** 2
    logits = np.zeros(A, np.float32)
    value, reward, hidden = C.c_double(), C.c_double(), C.c_uint64()
    lib().orc_e0_recurrent(C.byref(cfg), C.c_uint64(int(h_parent)), int(action), _p(logits, C.c_float),
                           C.byref(value), C.byref(reward), C.byref(hidden))
    return logits, value.value, reward.value, hidden.value


def pyset_order(keys):
    keys = np.ascontiguousarray(keys, dtype=np.int32)
    out = np.zeros_like(keys)
    lib().orc_pyset_order(_p(keys, C.c_int32), len(keys), _p(out, C.c_int32))
    return out


def check_win(board, n_in_row, r, c) -> bool:
    board = np.ascontiguousarray(board, dtype=np.int8)
    N = board.shape[0]
    return bool(lib().orc_check_win(_p(board, C.c_int8), N, int(n_in_row), int(r), int(c)))


def game_ended(board, n_in_row, last_move, move_count):
    """Returns +1/-1 (winner's stone), 0 (draw) or None -- game.py:60-63."""
    board = np.ascontiguousarray(board, dtype=np.int8)
    N = board.shape[0]
    w = lib().orc_game_ended(_p(board, C.c_int8), N, int(n_in_row), int(last_move), int(move_count))
    return None if w == 2 else int(w)


def selfplay_game(cfg: OrcConfig, gumbel):
    """gumbel: float64 [max_moves, A].  Returns dict(actions, values, policies, boards, winner)."""
    A = cfg.board_size ** 2
    gumbel = np.ascontiguousarray(gumbel, dtype=np.float64).reshape(-1, A)
    M = gumbel.shape[0]
    actions = np.zeros(M, np.int32)
    values = np.zeros(M, np.float64)
    policies = np.zeros((M, A), np.float64)
    boards = np.zeros((M, A), np.int8)
    winner = C.c_int32(2)
    T = lib().orc_selfplay_game(C.byref(cfg), _p(gumbel, C.c_double), M, _p(actions, C.c_int32),
                                _p(values, C.c_double), _p(policies, C.c_double), _p(boards, C.c_int8),
                                C.byref(winner))
    return dict(T=T, actions=actions[:T], values=values[:T], policies=policies[:T],
                boards=boards[:T].reshape(T, cfg.board_size, cfg.board_size),
                winner=None if winner.value == 2 else winner.value)


def final_rewards(T, winner):
    out = np.zeros(T, np.float32)
    lib().orc_final_rewards(int(T), int(winner), _p(out, C.c_float))
    return out


def n_step_returns(rewards, values, discount, n_steps):
    rewards = np.ascontiguousarray(rewards, dtype=np.float64)
    values = np.ascontiguousarray(values, dtype=np.float64)
    out = np.zeros(len(rewards), np.float32)
    lib().orc_n_step_returns(_p(rewards, C.c_double), _p(values, C.c_double), len(rewards), len(values),
                             C.c_double(discount), int(n_steps), _p(out, C.c_float))
    return out


def augment_batch(obs, act, pi, k, flip):
    """The D4 augmentation of calculate_loss (loss.py:37-51) as explicit index maps (small numpy loops).
    obs [B,U+1,3,N,N] and pi [B,U+1,A] move together: rot90 by k quarter turns (cell (r,c) ->
    (N-1-c, r) per turn, what torch.rot90 over the last two dims does), then a left-right flip.
    act [B,U] follows the reference's own formula (loss.py:46-51), which for k = 1, 3 is the OPPOSITE
    quarter turn of the one applied to the planes -- restated as written, including what it does to
    the -1 padding (floor division / non-negative remainder, as torch's // and % on integers)."""
    obs, act, pi = np.asarray(obs), np.asarray(act), np.asarray(pi)
    N = obs.shape[-1]
    dst = np.empty(N * N, np.int64)
    for r in range(N):
        for c in range(N):
            i, j = r, c
            for _ in range(int(k) % 4):
                i, j = N - 1 - j, i
            if flip:
                j = N - 1 - j
            dst[r * N + c] = i * N + j
    o = np.empty_like(obs).reshape(obs.shape[:-2] + (N * N,))
    o[..., dst] = obs.reshape(obs.shape[:-2] + (N * N,))
    p = np.empty_like(pi)
    p[..., dst] = pi
    a = act.astype(np.int64)
    rows, cols = a // N, a % N
    if k == 1:
        rows, cols = cols, N - 1 - rows
    elif k == 2:
        rows, cols = N - 1 - rows, N - 1 - cols
    elif k == 3:
        rows, cols = N - 1 - cols, rows
    if flip:
        cols = N - 1 - cols
    return o.reshape(obs.shape), (rows * N + cols).astype(act.dtype), p


class SumTree:
    """replay_buffer.SumTree restated over the C oracle (replay_buffer.py:4-41)."""

    def __init__(self, capacity):
        self.capacity = int(capacity)
        self.tree = np.zeros(2 * self.capacity - 1, np.float64)
        self.write_ptr = 0
        self.count = 0

    def update(self, tree_idx, priority):
        lib().orc_sumtree_update(_p(self.tree, C.c_double), C.c_int64(int(tree_idx)), C.c_double(float(priority)))

    def add(self, priority):
        self.update(self.write_ptr + self.capacity - 1, priority)
        self.write_ptr = (self.write_ptr + 1) % self.capacity
        if self.count < self.capacity:
            self.count += 1

    def get_leaf(self, value):
        return int(lib().orc_sumtree_get_leaf(_p(self.tree, C.c_double), C.c_int64(self.capacity), C.c_double(float(value))))

    def sample(self, B, u01, beta):
        u01 = np.ascontiguousarray(u01, dtype=np.float64)
        idx = np.zeros(B, np.int64)
        pr = np.zeros(B, np.float64)
        w = np.zeros(B, np.float32)
        lib().orc_per_sample(_p(self.tree, C.c_double), C.c_int64(self.capacity), C.c_int64(self.count), int(B),
                             _p(u01, C.c_double), C.c_double(beta), _p(idx, C.c_int64), _p(pr, C.c_double),
                             _p(w, C.c_float))
        return idx, pr, w

    def update_batch(self, tree_idx, priorities, max_priority):
        tree_idx = np.ascontiguousarray(tree_idx, dtype=np.int64)
        priorities = np.ascontiguousarray(priorities, dtype=np.float64)
        return float(lib().orc_per_update(_p(self.tree, C.c_double), len(tree_idx), _p(tree_idx, C.c_int64),
                                          _p(priorities, C.c_double), C.c_double(max_priority)))
