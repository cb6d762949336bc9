"""Host-side GomokuGame with the reference's board/move API (game.py:4-63), for callers that
hold a single game (tests, web UI, the batch-of-1 `search(game)` path).  The batched engine
keeps its own bitboard copy of every game on the device; this class is only the exchange format.
"""
from __future__ import annotations

import numpy as np

from .config import config

_DIRS = ((0, 1), (1, 0), (1, 1), (1, -1))


class GomokuGame:
    def __init__(self, board_size=None, n_in_row=None):
        self.board_size = config.BOARD_SIZE if board_size is None else board_size
        self.n_in_row = config.N_IN_ROW if n_in_row is None else n_in_row
        self.reset()

    def reset(self):
        n = self.board_size
        self.board = np.zeros((n, n), dtype=np.int8)
        self.current_player, self.last_move, self.move_count = 1, None, 0
        return self

    def get_board_state(self, player, last_move):
        """float32 [3,N,N]: stones of `player`, stones of the opponent, last-move one-hot."""
        planes = np.zeros((3, self.board_size, self.board_size), dtype=np.float32)
        planes[0][self.board == player] = 1.0
        planes[1][self.board == -player] = 1.0
        if last_move is not None:
            planes[2, last_move[0], last_move[1]] = 1.0
        return planes

    def get_valid_moves(self):
        rows, cols = np.nonzero(self.board == 0)
        return list(zip(rows, cols))

    def do_move(self, move_idx):
        r, c = divmod(int(move_idx), self.board_size)
        self.board[r, c] = self.current_player          # no legality check: overwrites (game.py:22)
        self.last_move = (r, c)
        self.current_player = -self.current_player
        self.move_count += 1

    def _longest_through(self, r, c, colour):
        """Length of the run of `colour` through (r, c) in each of the 4 directions, looking at most
        n_in_row + 1 cells each way (game.py:41-55) -- one fancy-indexed window per direction."""
        n, reach = self.board_size, self.n_in_row + 1
        k = np.arange(-reach, reach + 1)
        best = 0
        for dr, dc in _DIRS:
            rr, cc = r + k * dr, c + k * dc
            inside = (rr >= 0) & (rr < n) & (cc >= 0) & (cc < n)
            same = np.zeros(k.size, dtype=bool)
            same[inside] = self.board[rr[inside], cc[inside]] == colour
            gaps_before = np.flatnonzero(~same[:reach])
            gaps_after = np.flatnonzero(~same[reach + 1:])
            lo = gaps_before[-1] + 1 if gaps_before.size else 0
            hi = reach + 1 + (gaps_after[0] if gaps_after.size else reach)
            best = max(best, int(hi - lo))
        return best

    def check_win(self, move=None):
        where = self.last_move if move is None else move
        if where is None:
            return False
        colour = self.board[where[0], where[1]]
        return bool(colour != 0 and self._longest_through(where[0], where[1], colour) >= self.n_in_row)

    def get_game_ended(self):
        if self.check_win():
            return self.board[self.last_move[0], self.last_move[1]]
        if self.move_count >= self.board_size * self.board_size:
            return 0
        return None
