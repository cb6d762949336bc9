"""Generate golden vectors by importing the UNMODIFIED reference from /root/reference.

Runs only in the build container (the reference does not travel to the GPU box); the .npz
files it writes next to itself are committed and are what tests/ read.  Usage:

    python tests/golden/make_golden.py all          # every group (a few minutes)
    python tests/golden/make_golden.py search_az | search_mz | selfplay | game | per | tactics | network | augment

Nothing here is copied from the reference: the reference is imported and driven through its
public interface with (a) the E0 evaluator behind its queue protocol and (b) np.random.seed
so the Gumbel noise it draws is reproducible (RandomState(seed).gumbel(0, 1, A)).
"""
from __future__ import annotations

import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("GMZ_REFERENCE", "/root/reference")
sys.path.insert(0, HERE)
sys.dont_write_bytecode = True


def _import_reference(board_size, n_in_row=5):
    """config must be mutated before `game` is imported (game.py:5 binds defaults)."""
    sys.path.insert(0, os.path.join(HERE, "_stubs"))
    sys.path.insert(0, REF)
    from config import config
    config.BOARD_SIZE = board_size
    config.N_IN_ROW = n_in_row
    config.ACTION_SPACE_SIZE = board_size * board_size
    import game, mcts  # noqa
    return config, game, mcts


def _apply(config, p):
    config.NUM_SIMULATIONS = p["S"]
    config.NUM_TOP_ACTIONS = p["K"]
    config.C_VISIT = p.get("c_visit", 30)
    config.C_SCALE = p.get("c_scale", 1.0)
    config.DISCOUNT = p.get("discount", 0.997)
    config.VALUE_MINMAX_DELTA = p.get("delta", 1e-3)


def _random_position(N, n_moves, rs, game_mod, n_in_row):
    g = game_mod.GomokuGame(board_size=N, n_in_row=n_in_row)
    cells = rs.permutation(N * N)[:n_moves]
    for a in cells:
        g.do_move(int(a))
    return g


def _search_cases(N):
    A = N * N
    base = dict(S={6: 50, 9: 100, 15: 400, 19: 200}[N], K=16)      # 19x19 = the largest board the engine takes
    cases = []
    for n_moves in [0, 1, 2, 7, A // 3, A // 2, A - 20, A - 9, A - 3, A - 1]:
        cases.append(dict(base, n_moves=n_moves))
    for K in [1, 2, 4, 5, 8, 32]:
        cases.append(dict(base, K=K, n_moves=4))
    for S in [2, 15, 17, 33]:
        cases.append(dict(base, S=S, n_moves=3))
        cases.append(dict(base, S=S, n_moves=A - 6))
    cases.append(dict(base, n_moves=6, logit_div=4))
    cases.append(dict(base, n_moves=6, logit_div=2))
    cases.append(dict(base, n_moves=5, c_visit=50, c_scale=0.1))
    cases.append(dict(base, n_moves=5, discount=1.0, delta=0.01))
    cases.append(dict(base, n_moves=5, kind=1, const_value=0.5))       # MockModel
    cases.append(dict(base, n_moves=5, kind=1, const_value=1.5, const_reward=0.25))  # exercises the clip
    if N >= 15:
        cases = cases[:10] + cases[10:16:2] + cases[-6:]
    # dense (unquantised) logits / values: logit_div = 0 -- what a real network's outputs look like
    for n_moves in ([0, 5, A // 3, A - 9] if N < 15 else [0, 9, A // 2]):
        cases.append(dict(base, n_moves=n_moves, logit_div=0))
    # the production dtype: the evaluator returns np.float32 scalars like the reference's inference server
    # (workers.py:355,368) -> float32 value_sum / Q / MinMaxStats under NumPy >= 2
    f32 = [dict(base, n_moves=0), dict(base, n_moves=6), dict(base, n_moves=A // 2),
           dict(base, n_moves=0, logit_div=0), dict(base, n_moves=4, logit_div=0), dict(base, n_moves=A // 3, logit_div=0),
           dict(base, n_moves=A - 5, logit_div=0), dict(base, n_moves=3, logit_div=0, K=8),
           dict(base, n_moves=7, logit_div=0, S=33), dict(base, n_moves=5, logit_div=0, discount=1.0, delta=0.01),
           dict(base, n_moves=5, kind=1, const_value=0.5), dict(base, n_moves=5, kind=1, const_value=1.5, const_reward=0.25)]
    if N >= 15:
        f32 = f32[:2] + f32[3:7] + f32[-2:]
    cases += [dict(c, vdtype=1) for c in f32]
    return cases


def gen_search(mode_name, N):
    from e0_py import E0Queue
    config, game_mod, mcts = _import_reference(N)
    Engine = mcts.AlphaZeroMCTS if mode_name == "az" else mcts.MuZeroMCTS
    A = N * N
    rows = []
    for ci, p in enumerate(_search_cases(N)):
        _apply(config, p)
        seed = 1000 * N + ci
        rs = np.random.RandomState(seed)
        g = _random_position(N, p["n_moves"], rs, game_mod, 5)
        q = E0Queue(seed=seed, logit_div=p.get("logit_div", 16), kind=p.get("kind", 0),
                    const_value=p.get("const_value", 0.5), const_reward=p.get("const_reward", 0.0),
                    value_dtype=np.float32 if p.get("vdtype", 0) else None)
        q.set_action_space(A)
        eng = Engine(0, q, q)
        # capture the root Node and the per-evaluation leaf (action, depth)
        made = []
        OrigNode = mcts.Node

        class RecNode(OrigNode):
            def __init__(self, action=None, parent=None):
                super().__init__(action, parent)
                if parent is None:
                    made.append(self)
        mcts.Node = RecNode
        trace_a, trace_d = [], []
        orig_select = eng._select_leaf

        def traced(root, valid, mm, _o=orig_select):
            leaf, act = _o(root, valid, mm)
            d, n = 0, leaf
            while n.parent is not None:
                d += 1
                n = n.parent
            trace_a.append(int(act)); trace_d.append(d)
            return leaf, act
        eng._select_leaf = traced
        batch_sizes = []
        orig_rec = eng._remote_recurrent_inference_batch

        def rec_batch(h, acts, _o=orig_rec):
            batch_sizes.append(len(acts))
            return _o(h, acts)
        eng._remote_recurrent_inference_batch = rec_batch
        board0 = g.board.copy()
        np.random.seed(seed)
        policy, value, action = eng.search(g)
        mcts.Node = OrigNode
        assert np.array_equal(board0, g.board), "search mutated the game"
        gumbel = np.random.RandomState(seed).gumbel(0, 1, A)
        root = made[0]
        visits = np.zeros(A, np.int32)
        for a, ch in root.children.items():
            visits[int(a)] = ch.visit_count
        if mode_name == "mz":  # K selections per evaluation reach the same leaf: keep the first of each batch
            starts = np.concatenate([[0], np.cumsum(batch_sizes)[:-1]]).astype(int) if batch_sizes else []
            for s0, bs in zip(starts, batch_sizes):
                assert len(set(zip(trace_a[s0:s0 + bs], trace_d[s0:s0 + bs]))) == 1, "batch reached >1 leaf"
            ta, td = [trace_a[i] for i in starts], [trace_d[i] for i in starts]
        else:
            ta, td = trace_a, trace_d
        lm = -1 if g.last_move is None else int(g.last_move[0]) * N + int(g.last_move[1])
        rows.append(dict(
            params=np.array([N, 5, p["S"], p["K"], p.get("kind", 0), p.get("logit_div", 16), seed, p.get("vdtype", 0)], np.int64),
            value_is_f32=int(isinstance(value, np.float32)),
            fparams=np.array([p.get("c_visit", 30), p.get("c_scale", 1.0), p.get("delta", 1e-3),
                              p.get("discount", 0.997), p.get("const_value", 0.5), p.get("const_reward", 0.0)], np.float64),
            board=board0.astype(np.int8), player=int(g.current_player), last_move=lm, move_count=int(g.move_count),
            gumbel=gumbel, policy=np.asarray(policy, np.float64), value=float(value), action=int(action),
            visits=visits, root_n=int(root.visit_count), root_w=float(root.value_sum),
            n_initial=q.n_initial, n_recurrent=q.n_recurrent,
            leaf_actions=np.array(ta, np.int32), leaf_depths=np.array(td, np.int32),
            batch_sizes=np.array(batch_sizes, np.int32)))
        assert isinstance(value, np.float32) == bool(p.get("vdtype", 0)), type(value)
        print(f"[{mode_name} N={N}] case {ci}: S={p['S']} K={p['K']} moves={p['n_moves']} div={p.get('logit_div', 16)} "
              f"f32={p.get('vdtype', 0)} -> action {action} "
              f"value {float(value):+.6f} maxvisit {visits.max()} evals {q.n_initial}+{q.n_recurrent}", flush=True)
    out = {}
    for i, r in enumerate(rows):
        for k, v in r.items():
            out[f"c{i}_{k}"] = v
    out["n_cases"] = len(rows)
    np.savez_compressed(os.path.join(HERE, f"search_{mode_name}_{N}.npz"), **out)


def gen_selfplay(N, S, seed, mode_name="az", logit_div=16, vdtype=0):
    """One whole game through the reference's own universal_worker (workers.py:129-241).
    logit_div 0 = dense E0 logits; vdtype 1 = np.float32 evaluator values (the production dtype)."""
    from e0_py import E0Queue
    import queue
    config, game_mod, mcts = _import_reference(N)
    config.NUM_SIMULATIONS = S
    config.NUM_TOP_ACTIONS = 16
    config.MCTS_IMPLEMENTATION = "AlphaZero" if mode_name == "az" else "MuZero"
    import workers

    class Flag:
        def __init__(self): self.v = False
        def is_set(self): return self.v
        def set(self): self.v = True

    class Val:
        def __init__(self, v): self.value = v

    class Sink:
        def __init__(self, on_put=None): self.items, self.on_put = [], on_put
        def put(self, x):
            self.items.append(x)
            if self.on_put: self.on_put()
        def full(self): return False

    shutdown = Flag()
    data_q = Sink(on_put=shutdown.set)
    q = E0Queue(seed=seed, logit_div=logit_div, value_dtype=np.float32 if vdtype else None)
    q.set_action_space(N * N)
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    os.chdir(tmp)
    try:
        np.random.seed(seed)
        workers.universal_worker(0, Val(0), data_q, Sink(), Sink(), shutdown, q, q, Sink(), Sink(), Val(7),
                                 queue.Queue(), Flag())
    finally:
        os.chdir(cwd)
    rec, slices, version = data_q.items[0]
    T = len(rec.actions)
    U = config.NUM_UNROLL_STEPS
    out = dict(
        params=np.array([N, 5, S, 16, seed, U, config.N_STEPS, version, logit_div, vdtype], np.int64),
        discount=float(config.DISCOUNT),
        actions=np.array(rec.actions, np.int32),
        rewards=np.array(rec.rewards, np.float64),
        values_targets=np.array(rec.values, np.float64),
        policies=np.stack(rec.policies).astype(np.float64),
        boards=np.stack(rec.board_states).astype(np.int8),
        observations=np.stack(rec.observations).astype(np.float32),
        n_slices=len(slices),
        slice_obs=np.stack([s.observation for s in slices]),
        slice_act=np.stack([s.action_history for s in slices]),
        slice_rew=np.stack([s.reward_history for s in slices]),
        slice_pi=np.stack([s.policy_history for s in slices]),
        slice_val=np.stack([s.value_history for s in slices]),
    )
    out["slice_dtypes"] = np.array([str(slices[0].observation.dtype), str(slices[0].action_history.dtype),
                                    str(slices[0].reward_history.dtype), str(slices[0].policy_history.dtype),
                                    str(slices[0].value_history.dtype)])
    # the search values (bootstrap inputs) are not in the record: re-derive them by re-running
    # the reference search on every stored position with the same noise stream
    np.random.seed(seed)
    eng = (mcts.AlphaZeroMCTS if mode_name == "az" else mcts.MuZeroMCTS)(0, q, q)
    g = game_mod.GomokuGame(board_size=N, n_in_row=5)
    sv = []
    for t in range(T):
        pol, val, act = eng.search(g)
        assert act == rec.actions[t]
        sv.append(float(val))
        g.do_move(act)
    out["search_values"] = np.array(sv, np.float64)
    out["winner"] = int(g.get_game_ended())
    print(f"[selfplay {mode_name} N={N} S={S}] T={T} winner={out['winner']} slices={len(slices)}", flush=True)
    suffix = ("_dense" if logit_div == 0 else "") + ("_f32" if vdtype else "")
    np.savez_compressed(os.path.join(HERE, f"selfplay_{mode_name}_{N}_{S}{suffix}.npz"), **out)


def gen_game():
    """check_win / get_game_ended KATs (game.py:25-63): no upstream tests exist for these."""
    rows = []
    config, game_mod, _ = _import_reference(15)   # sizes are passed explicitly to GomokuGame below
    for N, nir in [(6, 5), (9, 5), (15, 5), (15, 6), (9, 4)]:
        rs = np.random.RandomState(7 * N + nir)
        for k in range(400):
            g = game_mod.GomokuGame(board_size=N, n_in_row=nir)
            fill = rs.uniform(0.2, 1.0)
            dens = rs.uniform(0.5, 0.9)   # biased colours make long runs likely
            b = np.where(rs.rand(N, N) < fill, np.where(rs.rand(N, N) < dens, 1, -1), 0).astype(np.int8)
            if k % 7 == 0:  # plant an exact run incl. overlines
                L = int(rs.randint(nir - 1, nir + 3)); r0 = int(rs.randint(0, N)); c0 = int(rs.randint(0, max(1, N - L)))
                b[r0, c0:c0 + L] = 1
            g.board = b
            occ = np.argwhere(b != 0)
            if len(occ) == 0:
                continue
            r, c = occ[rs.randint(len(occ))]
            g.last_move = (int(r), int(c))
            g.move_count = int((b != 0).sum())
            w = g.get_game_ended()
            rows.append((N, nir, b.reshape(-1).copy(), int(r) * N + int(c), g.move_count, 2 if w is None else int(w),
                         int(bool(g.check_win()))))
    out = dict(n=len(rows))
    out["N"] = np.array([r[0] for r in rows], np.int32)
    out["nir"] = np.array([r[1] for r in rows], np.int32)
    out["boards"] = np.array([np.pad(r[2], (0, 225 - len(r[2]))) for r in rows], np.int8)
    out["last"] = np.array([r[3] for r in rows], np.int32)
    out["move_count"] = np.array([r[4] for r in rows], np.int32)
    out["ended"] = np.array([r[5] for r in rows], np.int32)
    out["win"] = np.array([r[6] for r in rows], np.int32)
    print(f"[game] {len(rows)} positions, {int(out['win'].sum())} wins, {(out['ended'] == 0).sum()} draws", flush=True)
    np.savez_compressed(os.path.join(HERE, "game_kat.npz"), **out)


def gen_per():
    """SumTree / InMemoryReplayBuffer KATs (replay_buffer.py:4-106): no upstream tests exist."""
    sys.path.insert(0, REF)
    from config import config
    import replay_buffer as rb
    config.ENABLE_PER = True
    out = {}
    for ci, (cap, n_add, B, rounds) in enumerate([(8, 8, 4, 6), (37, 50, 8, 8), (1024, 700, 64, 6), (4096, 6000, 360, 5)]):
        np.random.seed(100 + ci)
        buf = rb.InMemoryReplayBuffer(cap)
        log_idx, log_w, log_td, log_tree, log_u = [], [], [], [], []
        rs = np.random.RandomState(55 + ci)
        for i in range(n_add):
            buf.add(i)
            if i % 5 == 0 and len(buf) >= 1:   # interleave priority updates so max_priority moves
                k = int(rs.randint(0, len(buf)))
                buf.update_priorities([k + cap - 1], rs.randn(1).astype(np.float32) * 3)
        out[f"p{ci}_tree_after_add"] = buf.sum_tree.tree.copy()
        out[f"p{ci}_state_after_add"] = np.array([buf.sum_tree.write_ptr, buf.sum_tree.count], np.int64)
        out[f"p{ci}_maxp_after_add"] = float(buf.max_priority)
        for r in range(rounds):
            st = np.random.get_state()
            u = np.random.random_sample(B)      # the B doubles np.random.uniform will consume
            np.random.set_state(st)
            batch, idx, w = buf.sample(B)
            td = rs.randn(B).astype(np.float32) * (2.0 if r % 2 else 0.05)
            buf.update_priorities(idx, td)
            log_u.append(u); log_idx.append(np.array(idx, np.int64)); log_w.append(w); log_td.append(td)
            log_tree.append(buf.sum_tree.tree.copy())
        out[f"p{ci}_params"] = np.array([cap, n_add, B, rounds], np.int64)
        out[f"p{ci}_u"] = np.stack(log_u); out[f"p{ci}_idx"] = np.stack(log_idx)
        out[f"p{ci}_w"] = np.stack(log_w); out[f"p{ci}_td"] = np.stack(log_td)
        out[f"p{ci}_tree"] = np.stack(log_tree)
        out[f"p{ci}_maxp"] = float(buf.max_priority)
        out[f"p{ci}_data"] = np.array([-1 if d is None else d for d in buf.data], np.int64)
        print(f"[per] case {ci}: cap={cap} adds={n_add} B={B} total={buf.sum_tree.total_priority():.6f}", flush=True)
    out["beta"] = float(config.PER_BETA); out["eps"] = float(config.PER_EPSILON); out["n_cases"] = 4
    np.savez_compressed(os.path.join(HERE, "per_kat.npz"), **out)


def gen_tactics():
    """find_winning_moves_rebuilt KATs (workers.py:49-123) incl. the boards of tests/test_winning_moves.py."""
    config, game_mod, _ = _import_reference(15)
    import workers
    rows = []
    rs = np.random.RandomState(3)
    for k in range(120):
        N = [9, 15][k % 2]
        fill = rs.uniform(0.05, 0.5)
        b = np.where(rs.rand(N, N) < fill, np.where(rs.rand(N, N) < 0.6, 1, -1), 0).astype(np.int8)
        for player in (1, -1):
            w = workers.find_winning_moves_rebuilt(b.copy(), player)
            cls = np.zeros(N * N, np.int8)
            for (r, c) in w["five"]: cls[r * N + c] = 1
            for (r, c) in w["open_four"]: cls[r * N + c] = 2
            for (r, c) in w["combo"]: cls[r * N + c] = 3
            rows.append((N, player, np.pad(b.reshape(-1), (0, 225 - N * N)), np.pad(cls, (0, 225 - N * N))))
    np.savez_compressed(os.path.join(HERE, "tactics_kat.npz"), n=len(rows),
                        N=np.array([r[0] for r in rows], np.int32), player=np.array([r[1] for r in rows], np.int32),
                        boards=np.array([r[2] for r in rows], np.int8), cls=np.array([r[3] for r in rows], np.int8))
    print(f"[tactics] {len(rows)} boards", flush=True)


def gen_network():
    """GomokuNetEZ forward KAT (network.py:109-152): a small seeded network's state_dict + outputs."""
    import torch
    config, _, _ = _import_reference(6)
    config.NUM_RES_BLOCKS, config.NUM_FILTERS, config.HEAD_HIDDEN_DIM = 2, 16, 8
    import network
    torch.manual_seed(0)
    net = network.GomokuNetEZ(config)
    with torch.no_grad():      # make BatchNorm statistics and the zero-initialised bn2 gains non-trivial
        for m in net.modules():
            if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm1d)):
                m.running_mean.normal_(0, 0.2); m.running_var.uniform_(0.5, 1.5); m.weight.normal_(1, 0.2); m.bias.normal_(0, 0.2)
    net.eval()
    obs = (torch.rand(5, 3, 6, 6) < 0.3).float()
    act = torch.tensor([[0], [7], [35], [12], [3]])
    p, v, h = net.initial_inference(obs)
    p2, v2, h2, r2 = net.recurrent_inference(h, act)
    out = {"sd_" + k: t.numpy() for k, t in net.state_dict().items()}
    out.update(obs=obs.numpy(), act=act.numpy(), p=p.numpy(), v=v.numpy(), h=h.numpy(), p2=p2.numpy(), v2=v2.numpy(),
               h2=h2.numpy(), r2=r2.numpy(), cfg=np.array([6, 2, 16, 8], np.int64))
    np.savez_compressed(os.path.join(HERE, "network_kat.npz"), **out)
    print(f"[network] {len(net.state_dict())} tensors, {sum(t.numel() for t in net.parameters())} params", flush=True)


def gen_augment():
    """The D4 augmentation block of calculate_loss (loss.py:37-51) run through the reference itself:
    np.random.randint / choice are pinned to each (k, flip), F.cross_entropy and the model hooks capture
    what the reference feeds the networks (augmented observations, policies, actions)."""
    config, game, mcts = _import_reference(6)
    import torch
    import loss as ref_loss
    config.NUM_UNROLL_STEPS, config.N_STEPS = 5, 5
    config.DEVICE = torch.device("cpu")
    sp = np.load(os.path.join(HERE, "selfplay_az_6_36.npz"))
    T = len(sp["slice_obs"])
    pick = np.array([0, 3, 7, T // 2, T - 7, T - 5, T - 3, T - 1])     # includes slices padded at the game's end
    obs_b, act_b = torch.from_numpy(sp["slice_obs"][pick]), torch.from_numpy(sp["slice_act"][pick])
    rew_b, pi_b, val_b = (torch.from_numpy(sp[k][pick]) for k in ("slice_rew", "slice_pi", "slice_val"))
    B, U1, A = pi_b.shape
    out = dict(obs=obs_b.numpy(), act=act_b.numpy(), rew=rew_b.numpy(), pi=pi_b.numpy(), val=val_b.numpy())

    class Capture:
        def __init__(self):
            self.obs, self.acts, self.ce = [], [], []
            self.projection_net = type("P", (), {"fc2": type("F", (), {"out_features": 4})()})()
        def train(self): pass
        def eval(self): pass
        def initial_inference(self, o):
            self.obs.append(("last", o.clone())); return None, torch.zeros(o.shape[0], 1), None
        def representation(self, o):
            self.obs.append(("repr", o.clone())); return torch.zeros(o.shape[0], 2, 6, 6, requires_grad=True)
        def prediction(self, h):
            return torch.zeros(h.shape[0], A), torch.zeros(h.shape[0], 3)
        def dynamics(self, h, a):
            self.acts.append(a.clone()); return h.clone(), torch.zeros(h.shape[0], 3)
        def project(self, h, with_grad=True):
            return torch.zeros(h.shape[0], 4)

    real_ce, real_randint, real_choice = ref_loss.F.cross_entropy, np.random.randint, np.random.choice
    for k in range(4):
        for flip in (False, True):
            cap = Capture()
            def fake_ce(inp, target, reduction="none", _cap=cap):
                _cap.ce.append(target.clone()); return torch.zeros(inp.shape[0])
            ref_loss.F.cross_entropy = fake_ce
            np.random.randint = lambda n, _k=k: _k
            np.random.choice = lambda seq, _f=flip: _f
            try:
                try:
                    ref_loss.calculate_loss(cap, cap, (obs_b, act_b, rew_b, pi_b.float(), val_b), torch.ones(B))
                except Exception as ex:                                  # everything we need is captured before the
                    print(f"[augment k={k} flip={flip}] stopped after capture: {type(ex).__name__}: {ex}", flush=True)   # loss tail
            finally:
                ref_loss.F.cross_entropy, np.random.randint, np.random.choice = real_ce, real_randint, real_choice
            tag = f"k{k}f{int(flip)}"
            obs_aug = np.zeros_like(out["obs"]); pi_aug = np.zeros((B, U1, A), np.float32)
            act_aug = np.full(out["act"].shape, -99, np.int64)
            assert cap.obs[0][0] == "last" and cap.obs[1][0] == "repr"
            obs_aug[:, -1] = cap.obs[0][1].numpy(); obs_aug[:, 0] = cap.obs[1][1].numpy()
            pol = [t for t in cap.ce if t.dim() == 2 and t.shape[1] == A]    # policy targets: step 0, then each unroll step
            pi_aug[:, 0] = pol[0].numpy()
            steps = [s for s in range(U1 - 1) if (out["act"][:, s] != -1).any()]
            reprs = [o for tag_, o in cap.obs[2:] if tag_ == "repr"]
            assert len(steps) == len(cap.acts) == len(reprs) == len(pol) - 1
            for s, a, o, pt in zip(steps, cap.acts, reprs, pol[1:]):
                m = out["act"][:, s] != -1
                act_aug[m, s] = a.numpy(); obs_aug[m, s + 1] = o.numpy(); pi_aug[m, s + 1] = pt.numpy()
            out[f"obs_{tag}"], out[f"pi_{tag}"], out[f"act_{tag}"] = obs_aug, pi_aug, act_aug
    np.savez_compressed(os.path.join(HERE, "augment_kat.npz"), **out)
    print(f"[augment] B={B} U={U1 - 1} A={A}: 8 symmetries", flush=True)


def _sub(*args):
    subprocess.check_call([sys.executable, os.path.abspath(__file__), *map(str, args)])


if __name__ == "__main__":
    cmd = sys.argv[1] if len(sys.argv) > 1 else "all"
    if cmd == "all":
        for N in (6, 9, 15, 19):
            _sub("search_az", N); _sub("search_mz", N)
        _sub("selfplay", 6, 36, 11, "az"); _sub("selfplay", 9, 100, 12, "az"); _sub("selfplay", 6, 50, 13, "mz")
        _sub("selfplay", 9, 64, 14, "az", 0, 1); _sub("selfplay", 6, 50, 15, "mz", 0, 1)
        _sub("game"); _sub("per"); _sub("tactics"); _sub("network"); _sub("augment")
    elif cmd in ("search_az", "search_mz"):
        gen_search(cmd[-2:], int(sys.argv[2]))
    elif cmd == "selfplay":
        gen_selfplay(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5] if len(sys.argv) > 5 else "az",
                     int(sys.argv[6]) if len(sys.argv) > 6 else 16, int(sys.argv[7]) if len(sys.argv) > 7 else 0)
    elif cmd == "game":
        gen_game()
    elif cmd == "per":
        gen_per()
    elif cmd == "tactics":
        gen_tactics()
    elif cmd == "network":
        gen_network()
    elif cmd == "augment":
        gen_augment()
    else:
        raise SystemExit(f"unknown group {cmd}")
