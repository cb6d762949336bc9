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

    def _run(self, r, c, dr, dc, colour):
        k, n = 0, self.board_size
        for i in range(1, self.n_in_row + 2):           # at most n_in_row + 1 stones each way
            rr, cc = r + i * dr, c + i * dc
            if 0 <= rr < n and 0 <= cc < n and self.board[rr, cc] == colour:
                k += 1
            else:
                break
        return k

    def check_win(self, move=None):
        if move is None:
            if self.last_move is None:
                return False
            move = self.last_move
        r, c = move
        colour = self.board[r, c]
        if colour == 0:
            return False
        return any(1 + self._run(r, c, dr, dc, colour) + self._run(r, c, -dr, -dc, colour) >= self.n_in_row
                   for dr, dc in _DIRS)

    def get_game_ended(self):
        if self.check_win():
            return self.board[self.last_move[0], self.last_move[1]]
        if self.move_count >= self.board_size * self.board_size:
            return 0
        return None
