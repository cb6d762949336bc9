"""Network evaluator on the device: folded/fused inference equals the plain module, weight hot-swap
keeps a captured CUDA graph valid, and a stepwise search driven by the network matches a search whose
evaluator outputs are replayed through the oracle-checked constant path."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _net(seed, n=9, blocks=2, ch=32):
    import torch
    from datou_gomoku_muzero_b200.config import Config
    from datou_gomoku_muzero_b200.network import GomokuNetEZ
    torch.manual_seed(seed)
    net = GomokuNetEZ(Config(BOARD_SIZE=n, ACTION_SPACE_SIZE=n * n, NUM_RES_BLOCKS=blocks, NUM_FILTERS=ch, HEAD_HIDDEN_DIM=16))
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.2); m.running_var.uniform_(0.5, 1.5); m.weight.normal_(1, 0.2); m.bias.normal_(0, 0.2)
    return net


def test_device_evaluator_and_weight_hot_swap():
    import torch
    from datou_gomoku_muzero_b200.network import DeviceEvaluator
    net_a, net_b = _net(0), _net(1)
    obs = (torch.rand(64, 3, 9, 9, device="cuda") < 0.3).float()
    ev = DeviceEvaluator(net_a, obs, dtype=torch.float32, graph=True)
    lg, v = ev(obs)
    p, val, _ = net_a.cuda().eval().initial_inference(obs)
    torch.testing.assert_close(lg, p, rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(v, val.reshape(-1), rtol=1e-3, atol=1e-3)
    ev.update_weights(net_b.state_dict())                   # ModelWeightsUpdate (workers.py:331-335)
    lg2, v2 = ev(obs)
    p2, val2, _ = net_b.cuda().eval().initial_inference(obs)
    torch.testing.assert_close(lg2, p2, rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(v2, val2.reshape(-1), rtol=1e-3, atol=1e-3)
    assert (p - p2).abs().max() > 1e-2


@pytest.mark.parametrize("accum", ["float32", "float64"])
def test_network_driven_search_replayed_through_the_oracle(accum):
    """Stepwise search with a real network (continuous float32 logits and values): record every
    (observation -> logits, value) the network produced, then run the CPU oracle's tree logic with a table
    evaluator returning exactly those numbers -> same visit counts, moves, root values.  accum = float32 is the
    production dtype (the reference's inference server returns np.float32 scalars, workers.py:355)."""
    import torch
    from datou_gomoku_muzero_b200.engine import SearchEngine
    from datou_gomoku_muzero_b200.network import DeviceEvaluator
    from oracle import oracle
    N, S, G = 9, 64, 12
    A = N * N
    rs = np.random.RandomState(5)
    boards = np.zeros((G, A), np.int8); players = np.ones(G, np.int8)
    last = np.full(G, -1, np.int32); mc = np.zeros(G, np.int32)
    for g in range(G):
        p = 1
        for a in rs.permutation(A)[: 3 * g]:
            boards[g, a] = p; last[g] = a; p = -p; mc[g] += 1
        players[g] = p
    eng = SearchEngine(G, board_size=N, num_simulations=S, accum_dtype=accum)
    eng.set_roots(boards, players, last, mc)
    ev = DeviceEvaluator(_net(3), eng.leaf_obs, dtype=torch.float32, graph=False)
    gum_h = rs.gumbel(0, 1, (G, A))
    gum = torch.from_numpy(gum_h).cuda()
    table = {}

    def record(obs, lg, v):
        o, l, vv = obs.cpu().numpy(), lg.cpu().numpy(), v.cpu().numpy()
        for g in range(G):                                  # the evaluator is a pure function of the observation
            key = o[g].tobytes()
            cur = (l[g].copy(), float(vv[g]))
            if key in table:
                assert np.array_equal(table[key][0], cur[0]) and table[key][1] == cur[1]
            table[key] = cur

    def net_values(v):
        return v if accum == "float32" else v.double()       # float64 mode: the same numbers handed over as Python floats
    obs = eng.root_obs(); lg, v = ev(obs); record(obs, lg, v); eng.root_expand(lg, net_values(v), gum)
    for _ in range(S - 1):
        obs = eng.select()
        lg, v = ev(obs)
        record(obs, lg, v)
        eng.expand_backup(lg, net_values(v))
    pol, val, act, vis = (t.cpu().numpy() for t in eng.finalize())
    assert (vis.sum(1) == S - 1).all() and np.allclose(pol.sum(1), 1.0) and (np.abs(val) <= 1).all()

    def table_eval(board, player, last_move):               # what game.get_board_state would build (game.py:12-17)
        o = np.zeros((3, N, N), np.float32)
        o[0] = board == player; o[1] = board == -player
        if last_move >= 0:
            o[2, last_move // N, last_move % N] = 1.0
        return table[o.tobytes()]
    oracle.set_eval_callback(table_eval)
    try:
        cfg = oracle.make_config(board_size=N, num_simulations=S, eval_kind=3, accum_dtype=int(accum == "float32"))
        for g in range(G):
            r = oracle.search(cfg, boards[g], players[g], last[g], mc[g], gum_h[g])
            assert np.array_equal(vis[g], r["visits"]), (accum, g)
            assert act[g] == r["action"] and val[g] == r["value"], (accum, g)
            np.testing.assert_allclose(pol[g], r["policy"], rtol=1e-5, atol=1e-12)
    finally:
        oracle.set_eval_callback(None)
