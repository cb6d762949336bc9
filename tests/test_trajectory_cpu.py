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
