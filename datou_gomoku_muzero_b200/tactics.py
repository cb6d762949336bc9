"""Tactical move classifier on the device (reference workers.py:49-123) and the missed-win statistics
built on it (workers.py:191-203)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check
from .config import config

FIVE, OPEN_FOUR, COMBO = 1, 2, 3


def classify_boards(boards, players, board_size=None, n_in_row=None, device=None):
    """boards int8 [B,N,N] or [B,A], players +-1 [B] -> int8 [B,A] classes (device tensor)."""
    if not torch.cuda.is_available():
        raise _lib.GmzError("classify_boards needs a CUDA device (there is no CPU fallback)")
    lib = _lib.load()
    dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    b = torch.as_tensor(boards).to(device=dev, dtype=torch.int8)
    B = b.shape[0]
    N = board_size or (b.shape[1] if b.dim() == 3 else int(round(b.shape[1] ** 0.5)))
    b = b.reshape(B, N * N).contiguous()
    p = torch.as_tensor(players).to(device=dev, dtype=torch.int8).reshape(B).contiguous()
    out = torch.empty((B, N * N), dtype=torch.int8, device=dev)
    check(lib.gmz_tactics_classify(C.c_void_p(b.data_ptr()), C.c_void_p(p.data_ptr()), B, N,
                                   int(config.N_IN_ROW if n_in_row is None else n_in_row), C.c_void_p(out.data_ptr()),
                                   C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "gmz_tactics_classify")
    return out


def find_winning_moves_rebuilt(board, player):
    """Same return shape as the reference: {'five': [(r,c)...], 'open_four': [...], 'combo': [...]}."""
    board = np.asarray(board)
    N = board.shape[0]
    cls = classify_boards(board[None], [player], N).cpu().numpy()[0]
    out = {"five": [], "open_four": [], "combo": []}
    for a in np.flatnonzero(cls):
        out[{FIVE: "five", OPEN_FOUR: "open_four", COMBO: "combo"}[int(cls[a])]].append((int(a) // N, int(a) % N))
    return out


def missed_win_stats(game_record, board_size=None):
    """(missed_fives, missed_totals) of one finished game, as workers.py:191-203 counts them."""
    T = len(game_record.actions)
    if T == 0:
        return 0, 0
    N = board_size or np.asarray(game_record.board_states[0]).shape[0]
    boards = np.stack([np.asarray(b, np.int8) for b in game_record.board_states[:T]])
    players = np.where(np.arange(T) % 2 == 0, 1, -1).astype(np.int8)
    cls = classify_boards(boards, players, N)
    acts = torch.as_tensor(np.asarray(game_record.actions[:T], dtype=np.int64), device=cls.device)
    any_win = (cls > 0).any(dim=1)
    played = cls.gather(1, acts.reshape(-1, 1)).reshape(-1)
    missed = any_win & (played == 0)
    has_five = (cls == FIVE).any(dim=1)
    return int((missed & has_five).sum().item()), int(missed.sum().item())
