"""Tactics classifier kernel vs KATs recorded from the reference's find_winning_moves_rebuilt
(workers.py:49-123), incl. the hand-made boards of the reference's tests/test_winning_moves.py."""
import os

import numpy as np
import pytest

from _golden_util import GOLDEN_DIR

pytestmark = pytest.mark.gpu


def test_tactics_matches_reference_kat():
    from datou_gomoku_muzero_b200.tactics import classify_boards
    z = np.load(os.path.join(GOLDEN_DIR, "tactics_kat.npz"))
    assert int(z["n"]) >= 200 and (z["cls"] == 1).sum() > 10 and (z["cls"] == 2).sum() > 0 and (z["cls"] == 3).sum() > 10
    for N in (9, 15):
        sel = np.flatnonzero(z["N"] == N)
        A = N * N
        got = classify_boards(z["boards"][sel][:, :A], z["player"][sel], N, n_in_row=5).cpu().numpy()
        assert np.array_equal(got, z["cls"][sel][:, :A]), N


def test_reference_unit_cases():
    """tests/test_winning_moves.py:30-84 restated: open-four ends, double three, four-three, double four."""
    from datou_gomoku_muzero_b200.tactics import find_winning_moves_rebuilt
    N = 15
    b = np.zeros((N, N), np.int8); b[7, 5:8] = 1                      # _OOO_ -> both ends make an open four
    w = find_winning_moves_rebuilt(b, 1)
    assert (7, 4) in w["open_four"] and (7, 8) in w["open_four"]
    b = np.zeros((N, N), np.int8); b[7, 6] = b[7, 8] = 1; b[6, 7] = b[8, 7] = 1   # double three at the centre
    assert (7, 7) in find_winning_moves_rebuilt(b, 1)["combo"]
    b = np.zeros((N, N), np.int8); b[7, 3:7] = 1; b[7, 2] = -1          # XOOOO_ -> five at the open end
    assert (7, 7) in find_winning_moves_rebuilt(b, 1)["five"]
    assert find_winning_moves_rebuilt(np.zeros((N, N), np.int8), 1) == {"five": [], "open_four": [], "combo": []}
