"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: game sharding, per-rank noise streams,
weight broadcast from the trainer rank, trajectory gather, max/sum-over-ranks aggregation."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from datou_gomoku_muzero_b200 import parallel
    from datou_gomoku_muzero_b200.config import Config
    from datou_gomoku_muzero_b200.network import GomokuNetEZ
    torch.manual_seed(100 + rank)                      # ranks start with DIFFERENT weights
    net = GomokuNetEZ(Config(BOARD_SIZE=6, ACTION_SPACE_SIZE=36, NUM_RES_BLOCKS=1, NUM_FILTERS=8, HEAD_HIDDEN_DIM=4))
    before = torch.cat([p.reshape(-1) for p in net.parameters()]).clone()
    parallel.broadcast_weights(net, src=0)
    after = torch.cat([p.reshape(-1) for p in net.parameters()])
    # the same through the persistent flat buffer: perturb rank 1, publish again, forward still works
    fw = parallel.FlatWeights(net)
    if rank == 1:
        with torch.no_grad():
            for p in net.parameters():
                p.add_(1.0)
    fw.broadcast(src=0)
    flat_after = torch.cat([p.reshape(-1) for p in net.parameters()])
    assert torch.equal(flat_after, after) and all(p.data_ptr() >= f.data_ptr() for p in net.parameters() for f in fw.flats[:1])
    net.eval()
    with torch.no_grad():
        out = net.initial_inference(torch.zeros(2, 3, 6, 6))
    assert out[0].shape == (2, 36) and torch.isfinite(out[0]).all()
    lo, hi = parallel.shard_games(9, world, rank)
    recs = [dict(game=g, length=3 + g, winner=1, actions=np.arange(3 + g)) for g in range(lo, hi)]
    got = parallel.gather_finished_games(recs, dst=0)
    mx = parallel.max_over_ranks(10.0 + rank)
    sm = parallel.sum_over_ranks(float(hi - lo))
    q.put((rank, before.sum().item(), after.sum().item(), (lo, hi), None if got is None else [(r["rank"], r["game"]) for r in got],
           mx, sm, parallel.rank_noise_seed(5, rank)))
    dist.destroy_process_group()


def test_two_rank_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, b0, a0, s0, g0, mx0, sm0, n0), (r1, b1, a1, s1, g1, mx1, sm1, n1) = res
    assert b0 != b1 and a0 == a1 == b0                 # rank 1 now holds rank 0's weights
    assert s0 == (0, 5) and s1 == (5, 9)               # contiguous shards covering all 9 games
    assert g1 is None and g0 == [(0, g) for g in range(0, 5)] + [(1, g) for g in range(5, 9)]
    assert mx0 == mx1 == 11.0 and sm0 == sm1 == 9.0
    assert n0 != n1


def test_shard_games_covers_everything():
    from datou_gomoku_muzero_b200.parallel import shard_games
    for total in (1, 7, 4096, 32768):
        for world in (1, 2, 3, 8):
            spans = [shard_games(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
