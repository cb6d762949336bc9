"""Surge re-analysis (reference workers.py:243-305): re-run the search on every stored position of
finished games with the latest model and rebuild the policy / value targets.

The reference walks one game at a time, one position at a time (`mcts_engine.search(temp_game)` per
stored board, workers.py:254-266).  Here all positions of all supplied games go through the engine in
batches of G -- the positions are independent roots (`board_states[i]`, player = +1 if i even else -1,
last move = actions[i-1], move_count = i).
"""
from __future__ import annotations

import numpy as np

from .config import config
from .trajectory import compute_n_step_returns


def positions_of(game_record, board_size):
    """The roots the reference rebuilds from one GameRecord (workers.py:258-263)."""
    T = len(game_record.actions)
    N = board_size
    boards = np.stack([np.asarray(b, dtype=np.int8).reshape(N * N) for b in game_record.board_states[:T]])
    players = np.where(np.arange(T) % 2 == 0, 1, -1).astype(np.int8)
    last = np.array([-1] + [int(a) for a in game_record.actions[:T - 1]], dtype=np.int32)
    return boards, players, last, np.arange(T, dtype=np.int32)


def reanalyse(game_records, batch_search, board_size=None, gumbel_fn=None):
    """game_records: list of GameRecord.  batch_search: an `AlphaZeroMCTS.for_engine(...)`-style object
    (search_batch over exactly G roots).  Returns, per game, (new_policies [T,A] float64,
    new_value_targets list[float]) -- what `finish_reanalysis_for_game` is handed (workers.py:292-294).
    gumbel_fn(n, A) supplies the noise rows (default: np.random.gumbel like each reference search)."""
    N = board_size or config.BOARD_SIZE
    A = N * N
    eng = batch_search._batch["engine"]
    G = eng.G
    pos = [positions_of(gr, N) for gr in game_records]
    boards = np.concatenate([p[0] for p in pos]); players = np.concatenate([p[1] for p in pos])
    last = np.concatenate([p[2] for p in pos]); mc = np.concatenate([p[3] for p in pos])
    n = len(boards)
    pol = np.zeros((n, A), np.float64); val = np.zeros(n, np.float64)
    for s in range(0, n, G):
        e = min(n, s + G)
        k = e - s
        b = np.zeros((G, A), np.int8); b[:k] = boards[s:e]
        b[k:, :] = 1                                      # padding roots: full boards -> inactive games
        pl = np.ones(G, np.int8); pl[:k] = players[s:e]
        lm = np.full(G, -1, np.int32); lm[:k] = last[s:e]
        m = np.zeros(G, np.int32); m[:k] = mc[s:e]
        g = gumbel_fn(G, A) if gumbel_fn else np.random.gumbel(0, 1, (G, A))
        p, v, _ = batch_search.search_batch(b, pl, lm, m, g)
        pol[s:e], val[s:e] = p[:k], v[:k]
    out, off = [], 0
    for gr in game_records:
        T = len(gr.actions)
        rewards = np.array(gr.rewards, dtype=np.float32)          # workers.py:291: float32 here, unlike self-play
        targets = compute_n_step_returns(rewards, [np.float64(x) for x in val[off:off + T]], config.DISCOUNT, config.N_STEPS)
        out.append((pol[off:off + T].copy(), targets))
        off += T
    return out
