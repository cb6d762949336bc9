"""MuZero-mode hidden-state pool kernels (gmz_hidden_gather / gmz_hidden_scatter) against plain torch
indexing, and the device search built on them: fused input vs recurrent_fn(hidden, actions), CUDA-graph
replay vs eager -- same visits, moves and values."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _engine(G, N=6, S=20):
    from datou_gomoku_muzero_b200.engine import SearchEngine
    return SearchEngine(G, board_size=N, num_simulations=S, mode="MuZero")


@pytest.mark.parametrize("dtype,C,E", [("bfloat16", 128, 16), ("float32", 8, 4), ("bfloat16", 4, 0), ("float32", 3, 1),
                                       ("bfloat16", 6, 2)])
def test_gather_scatter_match_torch_indexing(dtype, C, E):
    import torch
    from datou_gomoku_muzero_b200._lib import check
    dt = getattr(torch, dtype)
    G, S, nodes, A = 37, 20, 9, 36
    e = _engine(G, 6, S)
    g = torch.Generator(device="cuda").manual_seed(C * 7 + E)
    pool = torch.randn((G * nodes, A, C), device="cuda", generator=g).to(dt)
    node = torch.randint(0, nodes, (G,), device="cuda", generator=g)
    slot = (torch.arange(G, device="cuda") * S + node).int()
    slot[::5] = -1
    action = torch.randint(0, A, (G,), device="cuda", generator=g).int()
    action[slot < 0] = -1
    embed = torch.randn(E, device="cuda", generator=g).to(dt) if E else None
    x = torch.full((G, A, C + E), 7.0, device="cuda").to(dt)
    es = pool.element_size()
    check(e.lib.gmz_hidden_gather(pool.data_ptr(), slot.data_ptr(), action.data_ptr(), G, S, nodes, A, C * es,
                                  embed.data_ptr() if E else None, E * es, x.data_ptr(), e._stream()), "gather")
    rows = (torch.arange(G, device="cuda") * nodes + node).long()
    want = torch.zeros_like(x)
    live = slot >= 0
    want[live, :, :C] = pool[rows[live]]
    if E:
        want[live.nonzero().flatten(), action[live].long(), C:] = embed
    assert torch.equal(x.view(torch.uint8), want.view(torch.uint8))
    # scatter: rows of `h` land in the pool rows of the live slots, every other row is untouched
    h = torch.randn((G, A, C), device="cuda", generator=g).to(dt)
    before = pool.clone()
    child = (torch.arange(G, device="cuda") * S + (node + 1) % nodes).int()
    child[1::4] = -1
    check(e.lib.gmz_hidden_scatter(pool.data_ptr(), child.data_ptr(), G, S, nodes, A * C * es, h.data_ptr(), e._stream()),
          "scatter")
    crow = (torch.arange(G, device="cuda") * nodes + (node + 1) % nodes).long()
    want_pool = before.clone()
    want_pool[crow[child >= 0]] = h[child >= 0]
    assert torch.equal(pool.view(torch.uint8), want_pool.view(torch.uint8))


def test_gather_rejects_bad_sizes():
    import torch
    e = _engine(4)
    buf = torch.zeros(1024, dtype=torch.uint8, device="cuda")
    slot = torch.zeros(4, dtype=torch.int32, device="cuda")
    assert e.lib.gmz_hidden_gather(buf.data_ptr(), slot.data_ptr(), None, 4, 20, 4, 4, 6, None, 0, buf.data_ptr(), None) != 0
    assert b"multiples of 4" in e.lib.gmz_last_error()
    assert e.lib.gmz_hidden_gather(buf.data_ptr(), slot.data_ptr(), None, 4, 20, 4, 4, 8, None, 16, buf.data_ptr(), None) != 0
    assert e.lib.gmz_hidden_scatter(buf.data_ptr(), slot.data_ptr(), 4, 20, 4, 10, buf.data_ptr(), None) != 0


def _small_net(N):
    import torch
    from datou_gomoku_muzero_b200.config import Config
    from datou_gomoku_muzero_b200.network import GomokuNetEZ
    torch.manual_seed(3)
    cfg = Config(BOARD_SIZE=N, ACTION_SPACE_SIZE=N * N, NUM_RES_BLOCKS=2, NUM_FILTERS=32, HEAD_HIDDEN_DIM=16)
    net = GomokuNetEZ(cfg).cuda().eval()
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.2); m.running_var.uniform_(0.5, 1.5); m.weight.normal_(1, 0.2); m.bias.normal_(0, 0.2)
    return net


@pytest.mark.parametrize("graph", [False, True])
def test_network_search_fused_input_and_graph_replay_match_plain_path(graph):
    """A real (small) GomokuNetEZ in the tree: the gather-built concatenated input + CUDA-graph replay must
    reproduce the search that calls recurrent_fn(hidden, actions) step by step."""
    import torch
    from datou_gomoku_muzero_b200.muzero import FoldedRecurrentInference, MuZeroDeviceSearch
    from datou_gomoku_muzero_b200.network import FoldedInitialInference
    N, S, G = 6, 40, 24
    net = _small_net(N)
    fi, fr = FoldedInitialInference(net, torch.float32), FoldedRecurrentInference(net, torch.float32)

    def initial(obs):
        p, v, h = fi(obs.contiguous(memory_format=torch.channels_last))
        return p.float().contiguous(), v.reshape(-1).float(), h

    rs = np.random.RandomState(5)
    gum = torch.from_numpy(rs.gumbel(0, 1, (G, N * N))).cuda()
    out = []
    for fused in (False, True):
        eng = _engine(G, N, S)
        rf = fr if fused else (lambda h, a: fr(h, a))          # a bare callable: no forward_fused, plain gather
        mz = MuZeroDeviceSearch(eng, initial, rf, graph=graph and fused)
        assert mz.fused == fused
        n = mz.search(gum)
        assert (mz.graph is not None) == (graph and fused)
        pol, val, act, vis = (t.cpu().numpy().copy() for t in eng.finalize())
        out.append((n, pol, val, act, vis))
        # a second search on the same engine reuses pool + graph
        n2 = mz.search(gum)
        pol2, val2, act2, vis2 = (t.cpu().numpy() for t in eng.finalize())
        assert n2 == n and np.array_equal(vis2, vis) and np.array_equal(act2, act)
    (n0, p0, v0, a0, s0), (n1, p1, v1, a1, s1) = out
    assert n0 == n1 and np.array_equal(s0, s1) and np.array_equal(a0, a1)
    np.testing.assert_allclose(v0, v1, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(p0, p1, rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("graph", [False, True])
def test_e0_search_with_graph_matches_oracle(graph):
    """64-bit hash hidden states (8-byte rows: the narrow copy path), eager and graph-replayed, vs the oracle."""
    import torch
    from datou_gomoku_muzero_b200.muzero import MuZeroDeviceSearch, TorchE0
    from oracle import oracle
    N, S, G, seed = 9, 100, 40, 13
    A = N * N
    rs = np.random.RandomState(8)
    boards = np.zeros((G, A), np.int8); players = np.ones(G, np.int8)
    last = np.full(G, -1, np.int32); mc = np.zeros(G, np.int32)
    for g in range(G):
        k, p = int(rs.randint(0, A - 1)), 1
        for a in rs.permutation(A)[:k]:
            boards[g, a] = p; last[g] = a; p = -p
        players[g] = p; mc[g] = k
    gumbel = rs.gumbel(0, 1, (G, A))
    eng = _engine(G, N, S)
    eng.set_roots(boards, players, last, mc)
    e0 = TorchE0(N, seed=seed)
    mz = MuZeroDeviceSearch(eng, e0.initial, e0.recurrent, graph=graph)
    mz.search(torch.from_numpy(gumbel).cuda())
    pol, val, act, vis = (t.cpu().numpy() for t in eng.finalize())
    cfg = oracle.make_config(board_size=N, num_simulations=S, mode=1, eval_seed=seed)
    opol, oval, oact, ovis = oracle.search_batch(cfg, boards, players, last, mc, gumbel)
    assert np.array_equal(vis, ovis) and np.array_equal(act, oact) and np.array_equal(val, oval)
