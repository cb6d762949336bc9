"""2+ GPU check of the off-hot-path collectives (config 4): NCCL weight broadcast from the trainer rank
into every rank's DeviceEvaluator, per-rank self-play shards with distinct noise, trajectory gather.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/nccl_check.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from datou_gomoku_muzero_b200 import parallel
from datou_gomoku_muzero_b200.config import Config
from datou_gomoku_muzero_b200.engine import SearchEngine
from datou_gomoku_muzero_b200.network import DeviceEvaluator, GomokuNetEZ
from datou_gomoku_muzero_b200.selfplay import SelfPlayEngine
from datou_gomoku_muzero_b200.trajectory import TrajectoryStore

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
N, A = 9, 81
torch.manual_seed(100 + rank)                           # every rank starts with different weights
net = GomokuNetEZ(Config(BOARD_SIZE=N, ACTION_SPACE_SIZE=A, NUM_RES_BLOCKS=2, NUM_FILTERS=32, HEAD_HIDDEN_DIM=16)).to(dev)
lo, hi = parallel.shard_games(64, world, rank)
eng = SearchEngine(hi - lo, board_size=N, num_simulations=32, device=dev)
ev = DeviceEvaluator(net, eng.leaf_obs, dtype=torch.float32, graph=True)
obs = (torch.rand(hi - lo, 3, N, N, device=dev) < 0.3).float()
before = ev(obs)[0].clone()
t0 = time.perf_counter()
parallel.broadcast_weights(net, src=0); torch.cuda.synchronize()
bc_ms = (time.perf_counter() - t0) * 1e3
ev.update_weights(net.state_dict())                     # hot swap under the captured CUDA graph
after = ev(obs)[0].clone()
ref = [torch.empty_like(after) for _ in range(world)] if (hi - lo) * world == 64 else None
chk = torch.tensor([float(after.double().sum())], device=dev)
allc = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(allc, chk)
# sharded self-play with per-rank noise, trajectories gathered on rank 0
sp = SelfPlayEngine(eng, "e0", seed=5, noise_seed=parallel.rank_noise_seed(7, rank))
traj = TrajectoryStore(eng, extra_slots=32)
fin = []
for _ in range(4):
    sp.play(moves_per_game=12, traj=traj); fin += traj.harvest()
slim = [dict(game=lo + r["game"], length=r["length"], winner=r["winner"], actions=r["actions"]) for r in fin]
got = parallel.gather_finished_games(slim, dst=0)
total_moves = parallel.sum_over_ranks(eng.play_counters()[0], device=dev)
if rank == 0:
    print(f"world {world}: weight broadcast {bc_ms:.2f} ms; logits changed on rank0: {bool((before - after).abs().max() > 0)} (rank 0 keeps its own weights -> False expected)")
    print(f"gathered {len(got)} finished games from ranks {sorted(set(r['rank'] for r in got))}; total moves {int(total_moves)}")
    firsts = {}
    for r in got: firsts.setdefault(r["rank"], r)
    if world > 1:
        a0, a1 = firsts[0]["actions"], firsts[1]["actions"]
        print("different noise streams per rank ->", "different games" if len(a0) != len(a1) or (a0 != a1).any() else "IDENTICAL games (bug)")
else:
    assert (before - after).abs().max() > 0, "rank > 0 must have received rank 0's weights"
dist.barrier()
print(f"rank {rank}: OK")
dist.destroy_process_group()
