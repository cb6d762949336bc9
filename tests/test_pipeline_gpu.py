"""End-to-end on the device with a real (small) GomokuNetEZ: network-driven self-play -> trajectory
store -> device slice store -> PER sampling -> trainer batch tuple, i.e. universal_worker +
inference_server_worker + data_loader_worker (workers.py:129-439) without a host hop in the data path."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_network_selfplay_to_training_batch():
    import torch
    from datou_gomoku_muzero_b200.config import Config, config
    from datou_gomoku_muzero_b200.engine import SearchEngine
    from datou_gomoku_muzero_b200.network import DeviceEvaluator, GomokuNetEZ
    from datou_gomoku_muzero_b200.replay_buffer import InMemoryReplayBuffer
    from datou_gomoku_muzero_b200.selfplay import SelfPlayEngine
    from datou_gomoku_muzero_b200.tactics import missed_win_stats
    from datou_gomoku_muzero_b200.trajectory import DeviceSliceStore, TrajectoryStore, build_game_record
    N, S, G = 6, 16, 32
    A = N * N
    torch.manual_seed(0)
    net = GomokuNetEZ(Config(BOARD_SIZE=N, ACTION_SPACE_SIZE=A, NUM_RES_BLOCKS=2, NUM_FILTERS=16, HEAD_HIDDEN_DIM=8))
    eng = SearchEngine(G, board_size=N, num_simulations=S)
    ev = DeviceEvaluator(net, eng.leaf_obs, dtype=torch.float32, graph=True)
    sp = SelfPlayEngine(eng, ev, noise_seed=1)
    traj = TrajectoryStore(eng, extra_slots=96)
    store = DeviceSliceStore(traj)
    saved = (config.ENABLE_PER, config.N_IN_ROW)
    config.ENABLE_PER = True
    try:
        buf = InMemoryReplayBuffer(512)
        records = []
        for _ in range(A):
            sp.step(traj=traj)
            fin = traj.harvest(recycle=False)
            records += fin
            for st in store.ingest(fin):
                buf.add(st)                                   # the "slice" held by the buffer is (slot, t)
            if len(records) >= 6:
                break
        assert len(records) >= 6 and len(buf) == sum(r["length"] for r in records)
        for r in records:                                     # a finished game is a legal Gomoku game
            assert len(set(r["actions"].tolist())) == r["length"] and r["winner"] in (-1, 0, 1)
            np.testing.assert_allclose(r["policies"].sum(1), 1.0, atol=1e-9)
        batch, idx, w = buf.sample(16)
        obs, act, rew, pi, val = store.batch(batch)
        U = config.NUM_UNROLL_STEPS
        assert obs.shape == (16, U + 1, 3, N, N) and act.shape == (16, U) and pi.shape == (16, U + 1, A)
        assert obs.dtype == torch.float32 and pi.dtype == torch.float64 and val.dtype == torch.float32 and w.dtype == np.float32
        by_slot = {r["slot"]: r for r in records}
        for b, (slot, t) in enumerate(batch):                 # first unrolled action / policy are the game's own
            r = by_slot[slot]
            assert int(act[b, 0]) == int(r["actions"][t])
            assert np.array_equal(pi[b, 0].cpu().numpy(), r["policies"][t])
        buf.update_priorities(idx, np.abs(np.random.RandomState(0).randn(16)).astype(np.float32))
        assert abs(buf.sum_tree.total_priority() - float(buf.sum_tree.tree[511:].sum())) < 1e-9
        mf, mt = missed_win_stats(build_game_record(records[0]))
        assert 0 <= mf <= mt <= records[0]["length"]
    finally:
        config.ENABLE_PER, config.N_IN_ROW = saved
