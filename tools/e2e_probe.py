"""Where the end-to-end batch path spends its time: host time inside submit(), wait time inside result()."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from datou_gomoku_muzero_b200.engine import SearchEngine
from datou_gomoku_muzero_b200.mcts import PipelinedBatchSearch
G, depth, steps = 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 4, 40
engs = [SearchEngine(G, board_size=15, n_in_row=5, num_simulations=400, num_top_actions=16) for _ in range(depth)]
pipe = PipelinedBatchSearch(engs, evaluator="e0", eval_seed=bench.E0_SEED, logit_div=bench.LOGIT_DIV)
hb, hp, hl, hm = bench.staggered_positions(G, 0)
gum = [np.random.RandomState(i).gumbel(0, 1, (G, 225)) for i in range(3)]
for i in range(depth):
    pipe.result(pipe.submit(hb, hp, hl, hm, gum[i % 3]))
torch.cuda.synchronize()
ts, tr, infl = 0.0, 0.0, []
t0 = time.perf_counter()
for i in range(steps):
    a = time.perf_counter(); infl.append(pipe.submit(hb, hp, hl, hm, gum[i % 3])); ts += time.perf_counter() - a
    if len(infl) >= depth:
        a = time.perf_counter(); pipe.result(infl.pop(0)); tr += time.perf_counter() - a
while infl:
    a = time.perf_counter(); pipe.result(infl.pop(0)); tr += time.perf_counter() - a
torch.cuda.synchronize()
T = time.perf_counter() - t0
print(f"depth {depth}: {G*400*steps/T/1e6:.1f} M sims/s; per batch {T/steps*1e3:.2f} ms, host in submit {ts/steps*1e3:.2f} ms, waiting in result {tr/steps*1e3:.2f} ms")
