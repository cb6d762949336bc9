"""Development bench of the network-evaluator (E1) path: GomokuNetEZ 8x128 at 15x15, bf16,
CUDA-graph captured, driven by the stepwise kernels.  python tools/netbench.py [--games 4096] [--sims 40]"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from bench import staggered_positions, N, K_TOP, A
from datou_gomoku_muzero_b200.config import Config
from datou_gomoku_muzero_b200.engine import SearchEngine
from datou_gomoku_muzero_b200.network import GomokuNetEZ, DeviceEvaluator

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=4096)
ap.add_argument("--sims", type=int, default=40)
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--no-graph", action="store_true")
ap.add_argument("--no-fold", action="store_true")
ap.add_argument("--cudnn-benchmark", action="store_true")
args = ap.parse_args()
G, S = args.games, args.sims
torch.manual_seed(0)
cfg = Config(BOARD_SIZE=N, ACTION_SPACE_SIZE=A, NUM_RES_BLOCKS=8, NUM_FILTERS=128, HEAD_HIDDEN_DIM=64)
net = GomokuNetEZ(cfg)
eng = SearchEngine(G, board_size=N, num_simulations=S, num_top_actions=K_TOP)
eng.set_roots(*staggered_positions(G, 0))
dt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[args.dtype]
torch.backends.cudnn.benchmark = args.cudnn_benchmark
ev = DeviceEvaluator(net, eng.leaf_obs, dtype=dt, graph=not args.no_graph, folded=not args.no_fold)
print("folded:", ev.folded is not None, "fused cudnn ops:", getattr(ev.folded, "fused", None))
if ev.folded is not None:
    ref = DeviceEvaluator(net, eng.leaf_obs, dtype=torch.float32, graph=False, folded=False)
    eng.root_obs(); l1, v1 = ev(eng.leaf_obs); l1, v1 = l1.clone(), v1.clone(); l2, v2 = ref(eng.leaf_obs)
    print("folded vs plain fp32: max |dlogit| %.4f  max |dvalue| %.4f" % ((l1 - l2).abs().max().item(), (v1 - v2).abs().max().item()))
gum = torch.empty((G, A), dtype=torch.float64, device="cuda"); eng.fill_gumbel(gum, 1, 0)

def timeit(f, n):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n

for _ in range(3): ev(eng.leaf_obs)
t_net = timeit(lambda: ev(eng.leaf_obs), 10)
flop = 1.064e9 * G
print(f"net forward B={G} {args.dtype}: {t_net:.3f} ms  -> {G/t_net*1e3/1e3:.1f} k evals/s, {flop/t_net/1e9:.1f} TFLOP/s")

def search():
    lg, v = ev(eng.root_obs()); eng.root_expand(lg, v, gum)
    for _ in range(S - 1):
        lg, v = ev(eng.select()); eng.expand_backup(lg, v)
    eng.finalize(want_visits=False)
search()
t = timeit(search, 2)
print(f"stepwise search with net: {S} sims x {G} games: {t:.1f} ms -> {G*S/t/1e3:.3f} M sims/s ; per sim step {t/S:.3f} ms (net {t_net:.3f})")
t_sel = timeit(lambda: eng.select(), 20); 
print(f"k_select alone: {t_sel*1e3:.1f} us")
