"""Host post-processing of finished games (no GPU needed): final rewards, n-step value targets and
TrainingSlice cutting against what the reference's own universal_worker produced."""
import os

import numpy as np

from _golden_util import GOLDEN_DIR


import pytest


@pytest.mark.parametrize("name", ["selfplay_az_9_100", "selfplay_az_6_36", "selfplay_mz_6_50", "selfplay_az_9_64_dense_f32",
                                  "selfplay_mz_6_50_dense_f32"])
def test_game_record_and_slices_match_reference_golden(name):
    """Host post-processing (final rewards, n-step targets, slice cutting) against the GameRecord /
    TrainingSlices the reference's universal_worker produced (tests/golden/selfplay_*.npz)."""
    from datou_gomoku_muzero_b200.config import config
    from datou_gomoku_muzero_b200.trajectory import build_game_record, cut_training_slices
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    N, nir, S, K, seed, U, n_steps, version = (int(x) for x in z["params"][:8])
    saved = (config.DISCOUNT, config.N_STEPS, config.NUM_UNROLL_STEPS)
    config.DISCOUNT, config.N_STEPS, config.NUM_UNROLL_STEPS = float(z["discount"]), n_steps, U
    try:
        T = len(z["actions"])
        rec = dict(game=0, length=T, winner=int(z["winner"]), actions=z["actions"], values=z["search_values"],
                   policies=z["policies"], start_board=np.zeros((N, N), np.int8), start_player=1,
                   start_move_count=0, start_last_move=-1)
        gr = build_game_record(rec)
        assert np.array_equal(np.stack(gr.observations), z["observations"])
        assert np.array_equal(np.stack(gr.board_states), z["boards"])
        assert gr.actions == list(z["actions"]) and np.array_equal(np.array(gr.rewards), z["rewards"])
        assert np.array_equal(np.array(gr.values, np.float64), z["values_targets"])
        sl = cut_training_slices(gr)
        assert len(sl) == int(z["n_slices"])
        for name, key in (("observation", "slice_obs"), ("action_history", "slice_act"), ("reward_history", "slice_rew"),
                          ("policy_history", "slice_pi"), ("value_history", "slice_val")):
            got = np.stack([getattr(s, name) for s in sl])
            assert got.dtype == z[key].dtype and np.array_equal(got, z[key]), name
    finally:
        config.DISCOUNT, config.N_STEPS, config.NUM_UNROLL_STEPS = saved


def test_reanalysis_writeback_windows_match_the_reference_deque_logic():
    """db_manager.py:189-203: a deque of U + 1 policies / value targets slides over the game, zero-padded at the end;
    window t replaces policy_history / value_history of the stored slice of move t (db_manager.py:209-214)."""
    from collections import deque
    from datou_gomoku_muzero_b200.data_structures import TrainingSlice
    from datou_gomoku_muzero_b200.reanalysis import apply_writeback, writeback_windows
    rs = np.random.RandomState(3)
    for T, U, A in ((1, 5, 9), (4, 5, 9), (11, 3, 16), (7, 0, 4)):
        pols = [rs.dirichlet(np.ones(A)) for _ in range(T)]
        vals = [float(np.float32(x)) for x in rs.uniform(-1, 1, T)]
        k = U + 1
        pd, vd, exp_p, exp_v = deque(maxlen=k), deque(maxlen=k), [], []
        for i in range(T + k - 1):                                    # the reference's loop, restated
            pd.append(pols[i] if i < T else np.zeros_like(pols[0]))
            vd.append(vals[i] if i < T else 0.0)
            if i >= k - 1:
                exp_p.append(np.array(pd)); exp_v.append(np.array(vd, dtype=np.float32))
        pw, vw = writeback_windows(pols, vals, unroll_steps=U)
        assert pw.shape == (T, k, A) and pw.dtype == np.float64 and vw.shape == (T, k) and vw.dtype == np.float32
        assert np.array_equal(pw, np.array(exp_p)) and np.array_equal(vw, np.array(exp_v))
        slices = [TrainingSlice(np.zeros((k, 3, 2, 2), np.float32), np.zeros(U, np.int32), np.zeros(U, np.float32),
                                np.ones((k, A)), np.ones(k, np.float32)) for _ in range(T)]
        out = apply_writeback(slices, pols, vals, unroll_steps=U)
        assert all(np.array_equal(o.policy_history, exp_p[t]) and np.array_equal(o.value_history, exp_v[t]) for t, o in enumerate(out))
        assert all(o.observation is s.observation for o, s in zip(out, slices))
