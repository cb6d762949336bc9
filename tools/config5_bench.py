"""BASELINE configs[4]: PER replay sampling / priority updates at capacity 1M, B = 360, and Surge
re-analysis throughput (stored positions re-searched with the latest evaluator).
    python tools/config5_bench.py [--positions 65536]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from bench import staggered_positions, N, S, K_TOP, A, E0_SEED, LOGIT_DIV
from datou_gomoku_muzero_b200 import replay_buffer as rb
from datou_gomoku_muzero_b200.engine import SearchEngine
from datou_gomoku_muzero_b200.mcts import PipelinedBatchSearch

ap = argparse.ArgumentParser()
ap.add_argument("--positions", type=int, default=65536)
ap.add_argument("--cpu", action="store_true", help="also time the CPU oracle's SumTree (test infrastructure) for context")
args = ap.parse_args()
out = {}
# ---- PER: capacity 1M (config.py:59), B = 360 (config.py:56), priorities |N(0,1)| + 1e-6
cap, B = 1_000_000, 360
rs = np.random.RandomState(0)
tree = rb.SumTree(cap)
pri = np.abs(rs.randn(cap)) + 1e-6
t0 = time.perf_counter(); tree.add_many(pri); torch.cuda.synchronize(); fill_s = time.perf_counter() - t0
u = torch.from_numpy(rs.random_sample((200, B))).cuda()
idx = torch.empty(B, dtype=torch.int64, device="cuda"); pr = torch.empty(B, dtype=torch.float64, device="cuda")
w = torch.empty(B, dtype=torch.float32, device="cuda")
newp = torch.from_numpy(np.abs(rs.randn(200, B)) + 1e-6).cuda()
import ctypes as C
lib, st = tree.lib, tree._stream()
def step(i):
    lib.gmz_per_sample(C.c_void_p(tree.tree.data_ptr()), cap, cap, C.c_void_p(u[i].data_ptr()), B, 0.4,
                       C.c_void_p(idx.data_ptr()), C.c_void_p(pr.data_ptr()), C.c_void_p(w.data_ptr()), st)
    lib.gmz_per_update(C.c_void_p(tree.tree.data_ptr()), cap, C.c_void_p(idx.data_ptr()), C.c_void_p(newp[i].data_ptr()), B, st)
for i in range(10): step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(10, 200): step(i)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 190
out["per"] = {"capacity": cap, "batch": B, "us_per_sample_plus_update": ms * 1e3, "samples_per_sec": B / (ms * 1e-3),
              "fill_1M_ordered_adds_s": fill_s, "algorithmic_bytes_per_batch": B * (168 + 320)}
if args.cpu:
    from oracle import oracle as O
    ot = O.SumTree(cap); ot.tree[:] = tree.tree.cpu().numpy(); ot.count = cap
    uh, nh = u.cpu().numpy(), newp.cpu().numpy()
    t0 = time.perf_counter()
    for i in range(200):
        ii, pp, ww = ot.sample(B, uh[i], 0.4); ot.update_batch(ii, nh[i], 1.0)
    out["per"]["cpu_oracle_us_per_sample_plus_update"] = (time.perf_counter() - t0) / 200 * 1e6
# ---- Surge re-analysis: stored positions through the pipelined host-batch search (4096 per batch)
G = 4096
engs = [SearchEngine(G, board_size=N, num_simulations=S, num_top_actions=K_TOP) for _ in range(4)]
pipe = PipelinedBatchSearch(engs, evaluator="e0", eval_seed=E0_SEED + 1, logit_div=LOGIT_DIV)
hb, hp, hl, hm = staggered_positions(G, 3)
gum = np.random.RandomState(1).gumbel(0, 1, (G, A))
for _ in range(4): pipe.result(pipe.submit(hb, hp, hl, hm, gum))
nb = max(4, args.positions // G)
t0 = time.perf_counter(); inflight = []
for i in range(nb):
    inflight.append(pipe.submit(hb, hp, hl, hm, gum))
    if len(inflight) >= 4: pipe.result(inflight.pop(0))
while inflight: pipe.result(inflight.pop(0))
dt = time.perf_counter() - t0
out["reanalysis"] = {"positions": nb * G, "seconds": dt, "positions_per_sec": nb * G / dt, "sims_per_sec": nb * G * S / dt,
                     "seconds_per_1M_positions": 1e6 / (nb * G / dt)}
print(json.dumps(out))
