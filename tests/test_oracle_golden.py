"""Pins the C oracle (oracle/gmz_oracle.c) to the reference: every golden vector in
tests/golden/ was produced by importing the unmodified Python reference
(tests/golden/make_golden.py).  Integer outputs must be bit-exact; float64 outputs may
differ by the softmax's last-bit rounding only (torch's vectorised exp vs libm)."""
import os

import numpy as np
import pytest

from _golden_util import GOLDEN_DIR, load_search_cases
from oracle import oracle as O


def _cfg(c):
    return O.make_config(board_size=c["N"], n_in_row=c["n_in_row"], num_simulations=c["S"], num_top_actions=c["K"],
                         mode=0 if c["mode"] == "az" else 1, eval_kind=c["kind"], logit_div=c["logit_div"],
                         c_visit=c["c_visit"], c_scale=c["c_scale"], minmax_delta=c["delta"], discount=c["discount"],
                         const_value=c["const_value"], const_reward=c["const_reward"], eval_seed=c["seed"],
                         accum_dtype=c["vdtype"])


@pytest.mark.parametrize("mode", ["az", "mz"])
@pytest.mark.parametrize("N", [6, 9, 15, 19])
def test_search_matches_reference(mode, N):
    cases = load_search_cases(mode, N)
    assert len(cases) >= 19
    assert sum(c["logit_div"] == 0 for c in cases) >= 7 and sum(c["vdtype"] == 1 for c in cases) >= 8   # dense / float32 cases
    for c in cases:
        r = O.search(_cfg(c), c["board"], c["player"], c["last_move"], c["move_count"], c["gumbel"], trace=True)
        tag = f"{mode} N={N} case {c['idx']} div={c['logit_div']} f32={c['vdtype']}"
        assert r["all_visited"] == 0, tag
        assert int(c["value_is_f32"]) == c["vdtype"], tag
        assert np.array_equal(r["visits"], c["visits"]), tag
        assert r["action"] == c["action"], tag
        assert np.array_equal(r["leaf_actions"], c["leaf_actions"]), tag
        assert np.array_equal(r["leaf_depths"], c["leaf_depths"]), tag
        assert r["value"] == c["value"], tag                      # float64 / float32 accumulation is bit-exact
        np.testing.assert_allclose(r["policy"], c["policy"], rtol=1e-12, atol=1e-15, err_msg=tag)
        assert abs(r["policy"].sum() - 1.0) < 1e-9
        if mode == "az":
            assert r["n_evals"] + 1 == c["n_initial"] == c["S"]   # tests/test_mcts_logic.py:116-125
        else:
            assert c["n_initial"] == 1 and r["n_evals"] == c["n_recurrent"]   # tests/test_mcts_logic.py:127-136


def test_e0_python_and_c_agree():
    import e0_py
    rs = np.random.RandomState(0)
    for N in (6, 9, 15, 19):
        A = N * N
        for trial in range(20):
            board = rs.randint(-1, 2, size=A).astype(np.int8)
            player = int(rs.choice([-1, 1]))
            last = int(rs.randint(-1, A))
            seed = int(rs.randint(0, 2**31))
            div = int(rs.choice([0, 2, 4, 16]))
            cfg = O.make_config(board_size=N, eval_seed=seed, logit_div=div)
            lg, v, h = O.e0_initial(cfg, board, player, last)
            obs = np.zeros((3, N, N), np.float32)
            obs[0] = (board.reshape(N, N) == player); obs[1] = (board.reshape(N, N) == -player)
            if last >= 0:
                obs[2, last // N, last % N] = 1
            hp = e0_py.hash_obs(obs, seed)
            lp, vp = e0_py.heads(hp, A, div)
            assert hp == h and vp == v and np.array_equal(lp, lg)
            a = int(rs.randint(0, A))
            lg2, v2, r2, h2 = O.e0_recurrent(cfg, h, a)
            hc = e0_py.child_hidden(hp, a)
            lp2, vp2 = e0_py.heads(hc, A, div)
            assert hc == h2 and vp2 == v2 and e0_py.reward_of(hc, div) == r2 and np.array_equal(lp2, lg2)
            assert np.float32(vp) == vp and np.float32(r2) == r2        # exactly representable in float32


def test_pyset_order_matches_cpython():
    """mcts.py:274-275 scans a set of np.int64 actions; the oracle restates CPython's set layout."""
    rs = np.random.RandomState(1)
    for trial in range(300):
        A = int(rs.choice([36, 81, 225, 361]))
        n = int(rs.randint(1, A + 1))
        keys = np.sort(rs.choice(A, size=n, replace=False))
        expect = list({np.int64(k) for k in keys})
        got = O.pyset_order(keys)
        assert [int(x) for x in expect] == [int(x) for x in got], (A, n)


def test_game_kat():
    z = np.load(os.path.join(GOLDEN_DIR, "game_kat.npz"))
    n = int(z["n"])
    assert n >= 1500 and z["win"].sum() > 100 and (z["ended"] == 0).sum() > 0
    for i in range(n):
        N, nir = int(z["N"][i]), int(z["nir"][i])
        b = z["boards"][i][:N * N].reshape(N, N)
        last = int(z["last"][i])
        assert O.check_win(b, nir, last // N, last % N) == bool(z["win"][i])
        w = O.game_ended(b, nir, last, int(z["move_count"][i]))
        assert (2 if w is None else w) == int(z["ended"][i])


def test_per_kat():
    z = np.load(os.path.join(GOLDEN_DIR, "per_kat.npz"))
    beta, eps = float(z["beta"]), float(z["eps"])
    for ci in range(int(z["n_cases"])):
        cap, n_add, B, rounds = (int(x) for x in z[f"p{ci}_params"])
        tree = O.SumTree(cap)
        tree.tree[:] = z[f"p{ci}_tree_after_add"]
        tree.write_ptr, tree.count = (int(x) for x in z[f"p{ci}_state_after_add"])
        maxp = float(z[f"p{ci}_maxp_after_add"])
        for r in range(rounds):
            idx, pr, w = tree.sample(B, z[f"p{ci}_u"][r], beta)
            assert np.array_equal(idx, z[f"p{ci}_idx"][r]), (ci, r)
            np.testing.assert_allclose(w, z[f"p{ci}_w"][r], rtol=1e-6)
            pri = (np.abs(z[f"p{ci}_td"][r]) + eps)            # float32, as the reference computes it
            assert pri.dtype == np.float32
            maxp = tree.update_batch(idx, pri.astype(np.float64), maxp)
            assert np.array_equal(tree.tree, z[f"p{ci}_tree"][r]), (ci, r)   # bit-exact float64 tree
        assert maxp == float(z[f"p{ci}_maxp"])


def test_sumtree_add_sequence():
    """SumTree.add ring semantics (replay_buffer.py:21-25) incl. wrap-around."""
    t = O.SumTree(5)
    for i in range(12):
        t.add(float(i + 1))
    assert t.count == 5 and t.write_ptr == 2
    assert t.tree[0] == pytest.approx(sum([11, 12, 8, 9, 10]))


@pytest.mark.parametrize("name,mode", [("selfplay_az_6_36", 0), ("selfplay_az_9_100", 0), ("selfplay_mz_6_50", 1),
                                       ("selfplay_az_9_64_dense_f32", 0), ("selfplay_mz_6_50_dense_f32", 1)])
def test_selfplay_game_matches_reference(name, mode):
    """Whole game through the reference's universal_worker vs the oracle's game loop."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    N, nir, S, K, seed, U, n_steps, version, logit_div, vdtype = (int(x) for x in z["params"])
    A = N * N
    cfg = O.make_config(board_size=N, n_in_row=nir, num_simulations=S, num_top_actions=K, mode=mode, eval_seed=seed,
                        logit_div=logit_div, accum_dtype=vdtype)
    gumbel = np.random.RandomState(seed).gumbel(0, 1, (A + 1, A))
    g = O.selfplay_game(cfg, gumbel)
    T = len(z["actions"])
    assert g["T"] == T and np.array_equal(g["actions"], z["actions"])
    assert g["winner"] == int(z["winner"])
    assert np.array_equal(g["boards"], z["boards"])
    assert np.array_equal(g["values"], z["search_values"])
    np.testing.assert_allclose(g["policies"], z["policies"], rtol=1e-12, atol=1e-15)
    rew = O.final_rewards(T, g["winner"])
    assert np.array_equal(rew.astype(np.float64), z["rewards"])
    vt = O.n_step_returns(rew.astype(np.float64), g["values"], float(z["discount"]), n_steps)
    assert np.array_equal(vt, z["values_targets"].astype(np.float32)), np.abs(vt - z["values_targets"]).max()


def test_augmentation_restatement_matches_reference_loss_path():
    """O.augment_batch vs the tensors the reference's calculate_loss fed its (capturing) networks for
    every (k, flip) -- loss.py:37-51, golden made by tests/golden/make_golden.py augment."""
    g = np.load(os.path.join(GOLDEN_DIR, "augment_kat.npz"))
    valid = g["act"] != -1
    for k in range(4):
        for f in (0, 1):
            o, a, p = O.augment_batch(g["obs"], g["act"], g["pi"], k, bool(f))
            tag = f"k{k}f{f}"
            assert np.array_equal(o, g["obs_" + tag]) and np.array_equal(p.astype(np.float32), g["pi_" + tag])
            assert np.array_equal(a[valid], g["act_" + tag][valid])
    # the reference's action formula is the inverse quarter turn of its plane rotation for k = 1, 3
    N = g["obs"].shape[-1]
    one = np.zeros((1, 1, 3, N, N), np.float32); one[0, 0, 2, 1, 2] = 1.0
    o, a, _ = O.augment_batch(one, np.array([[1 * N + 2]], np.int32), np.zeros((1, 1, N * N)), 1, False)
    assert int(np.argmax(o[0, 0, 2])) != int(a[0, 0]) and int(a[0, 0]) == 2 * N + (N - 1 - 1)
