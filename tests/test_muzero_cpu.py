"""Host-side MuZero logic: the number of evaluations per search follows the reference's halving state
machine (mcts.py:158-181 driven with sim_count += len(selected), mcts.py:346)."""
from _golden_util import load_search_cases


def test_evals_per_search_matches_reference_request_counts():
    from datou_gomoku_muzero_b200.muzero import evals_per_search
    n = 0
    for N in (6, 9, 15, 19):
        for c in load_search_cases("mz", N):
            n_valid = int((c["board"] == 0).sum())
            assert evals_per_search(c["S"], c["K"], min(c["K"], n_valid)) == c["n_recurrent"], (N, c["idx"])
            n += 1
    assert n > 60


def test_torch_e0_matches_python_e0():
    """The torch-integer E0 (device evaluator of the MuZero path) computes the same integers as e0_py."""
    import numpy as np
    import torch
    import e0_py
    from datou_gomoku_muzero_b200.muzero import TorchE0
    rs = np.random.RandomState(2)
    for N in (6, 9, 15, 19):
        A = N * N
        B = 12
        obs = np.zeros((B, 3, N, N), np.float32)
        for b in range(B):
            cells = rs.randint(-1, 2, size=(N, N))
            obs[b, 0] = cells == 1; obs[b, 1] = cells == -1
            if b % 3:
                a = rs.randint(A); obs[b, 2, a // N, a % N] = 1
        if N == 15:
            obs[0, 0].reshape(-1)[63] = 1; obs[0, 1].reshape(-1)[63] = 0      # exercise the sign bit of a word
        seed, div = int(rs.randint(1 << 30)), [0, 4, 16, 2][[6, 9, 15, 19].index(N)]
        e0 = TorchE0(N, seed=seed, logit_div=div, device="cpu")
        lg, v, h = e0.initial(torch.from_numpy(obs))
        acts = torch.from_numpy(rs.randint(0, A, size=B))
        lg2, v2, r2, h2 = e0.recurrent(h, acts)
        for b in range(B):
            hp = e0_py.hash_obs(obs[b], seed)
            assert (int(h[b]) & ((1 << 64) - 1)) == hp
            l_ref, v_ref = e0_py.heads(hp, A, div)
            assert np.array_equal(lg[b].numpy(), l_ref) and float(v[b]) == v_ref
            hc = e0_py.child_hidden(hp, int(acts[b]))
            assert (int(h2[b]) & ((1 << 64) - 1)) == hc
            l_ref2, v_ref2 = e0_py.heads(hc, A, div)
            assert np.array_equal(lg2[b].numpy(), l_ref2) and float(v2[b]) == v_ref2 and float(r2[b]) == e0_py.reward_of(hc, div)
