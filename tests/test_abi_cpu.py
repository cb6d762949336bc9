"""CPU-side checks of the boundary: the C-ABI library loads without a GPU and exports every
symbol include/gmz.h declares; host-side contract types and the game API behave like the
reference's; the product never imports the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "gmz.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gmz_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from datou_gomoku_muzero_b200 import _build, _lib
    _build.build()
    names = _declared()
    assert len(names) >= 20
    lib = ctypes.CDLL(_lib.SO)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gmz.h but not exported by libgmz.so"
    assert set(_lib.SIGNATURES) == set(names), "ctypes SIGNATURES out of sync with include/gmz.h"
    loaded = _lib.load()
    assert loaded.gmz_version() == 100


def test_config_validation_without_gpu():
    from datou_gomoku_muzero_b200 import _lib
    lib = _lib.load()
    ok = _lib.GmzConfig(15, 5, 400, 16, 0, 4096, 0, 0, 30.0, 1.0, 1e-3, 0.997)
    nbytes = lib.gmz_workspace_bytes(ctypes.byref(ok))
    assert 2.4e9 < nbytes < 3.0e9          # 4096 games x 400 nodes x (1 KiB logits + 512 B children + headers)
    for bad in (_lib.GmzConfig(20, 5, 400, 16, 0, 1, 0, 0, 30.0, 1.0, 1e-3, 0.997),      # board too large
                _lib.GmzConfig(15, 5, 0, 16, 0, 1, 0, 0, 30.0, 1.0, 1e-3, 0.997),        # no simulations
                _lib.GmzConfig(15, 5, 400, 33, 0, 1, 0, 0, 30.0, 1.0, 1e-3, 0.997),      # too many top actions
                _lib.GmzConfig(15, 5, 400, 16, 7, 1, 0, 0, 30.0, 1.0, 1e-3, 0.997)):     # unknown mode
        assert lib.gmz_workspace_bytes(ctypes.byref(bad)) == 0
        assert lib.gmz_last_error()
    assert lib.gmz_set_roots(None, None, None, None, None, None) != 0     # null engine is an error, not a crash


def test_engine_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from datou_gomoku_muzero_b200 import _lib
    from datou_gomoku_muzero_b200.engine import SearchEngine
    from datou_gomoku_muzero_b200.replay_buffer import InMemoryReplayBuffer
    with pytest.raises(_lib.GmzError):
        SearchEngine(1)
    with pytest.raises(_lib.GmzError):
        InMemoryReplayBuffer(8)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "datou_gomoku_muzero_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower(), f"{f} mentions the oracle"


def test_game_api_matches_reference_kat():
    """GomokuGame host class vs the reference KATs (game.py:20-63)."""
    from datou_gomoku_muzero_b200.game import GomokuGame
    z = np.load(os.path.join(ROOT, "tests", "golden", "game_kat.npz"))
    for i in range(0, int(z["n"]), 3):
        N, nir = int(z["N"][i]), int(z["nir"][i])
        g = GomokuGame(board_size=N, n_in_row=nir)
        g.board = z["boards"][i][:N * N].reshape(N, N).copy()
        last = int(z["last"][i])
        g.last_move, g.move_count = (last // N, last % N), int(z["move_count"][i])
        w = g.get_game_ended()
        assert (2 if w is None else int(w)) == int(z["ended"][i])
        assert bool(g.check_win()) == bool(z["win"][i])
    g = GomokuGame(board_size=6)
    assert g.get_board_state(1, None).shape == (3, 6, 6) and g.get_board_state(1, None).dtype == np.float32
    g.do_move(7); g.do_move(7)                       # overwrite is allowed (game.py:22)
    assert g.board[1, 1] == -1 and g.move_count == 2 and g.current_player == 1 and g.last_move == (1, 1)
    assert len(g.get_valid_moves()) == 35


def test_contract_types():
    from datou_gomoku_muzero_b200 import GameRecord, TrainingSlice
    assert GameRecord._fields == ("observations", "actions", "rewards", "policies", "values", "board_states")
    assert TrainingSlice._fields == ("observation", "action_history", "reward_history", "policy_history", "value_history")
