"""BASELINE configs[2]: MuZero-mode 15x15, 400 simulations, G games per GPU, real dynamics network in
the tree (GomokuNetEZ 8x128, bf16, folded + fused).  python tools/mzbench.py [--games 4096]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from bench import staggered_positions, N, S, K_TOP, A
from datou_gomoku_muzero_b200.config import Config
from datou_gomoku_muzero_b200.engine import SearchEngine
from datou_gomoku_muzero_b200.muzero import FoldedRecurrentInference, MuZeroDeviceSearch, evals_per_search
from datou_gomoku_muzero_b200.network import FoldedInitialInference, GomokuNetEZ

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=4096)
ap.add_argument("--searches", type=int, default=2)
ap.add_argument("--graph", type=int, default=1, help="capture the simulation step in a CUDA graph")
ap.add_argument("--profile", default="", help="write a per-kernel GPU-time table (torch profiler) of one search to this file")
args = ap.parse_args()
G = args.games
torch.manual_seed(0); torch.backends.cudnn.benchmark = True
cfg = Config(BOARD_SIZE=N, ACTION_SPACE_SIZE=A, NUM_RES_BLOCKS=8, NUM_FILTERS=128, HEAD_HIDDEN_DIM=64)
net = GomokuNetEZ(cfg).cuda().eval()
fi, fr = FoldedInitialInference(net, torch.bfloat16), FoldedRecurrentInference(net, torch.bfloat16)

def initial(obs):
    p, v, h = fi(obs.to(torch.bfloat16).contiguous(memory_format=torch.channels_last))
    return p.float().contiguous(), v.reshape(-1).float(), h

eng = SearchEngine(G, board_size=N, num_simulations=S, num_top_actions=K_TOP, mode="MuZero")
eng.set_roots(*staggered_positions(G, 0))
evals = evals_per_search(S, K_TOP, K_TOP)
mz = MuZeroDeviceSearch(eng, initial, fr, nodes_per_game=evals + 2, graph=bool(args.graph))
gum = torch.empty((G, A), dtype=torch.float64, device="cuda"); eng.fill_gumbel(gum, 1, 0)
mz.search(gum, max_steps=evals); eng.finalize(want_visits=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.searches + 1)]
marks[0].record()
for i in range(args.searches):
    steps = mz.search(gum, max_steps=evals); eng.finalize(want_visits=False)
    marks[i + 1].record()
torch.cuda.synchronize()
per_search = [marks[i].elapsed_time(marks[i + 1]) for i in range(args.searches)]
ms = sum(per_search) / args.searches
tf = (1.141e9 * steps + 1.064e9) * G / (ms * 1e-3) / 1e12
print(json.dumps({"config": "MuZero-mode 15x15, 400 sims, %d games, GomokuNetEZ 8x128 bf16 in-tree dynamics" % G,
                  "ms_per_search": ms, "recurrent_evals_per_search": steps, "sims_per_sec": G * S / (ms * 1e-3),
                  "moves_per_sec": G / (ms * 1e-3), "distinct_evals_per_sec": G * (steps + 1) / (ms * 1e-3),
                  "tensor_tflops": tf, "ms_each": [round(x, 1) for x in per_search], "graph": bool(args.graph), "hidden_pool_gb": mz.pool.numel() * mz.pool.element_size() / 1e9}))

if args.profile:
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        mz.search(gum, max_steps=evals); eng.finalize(want_visits=False)
        torch.cuda.synchronize()
    with open(args.profile, "w") as f:
        f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=90))
