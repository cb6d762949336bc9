"""Run in a subprocess with GMZ_LIB pointing at the -DGMZ_VERIFY_FAST build of the library: searches over
several shapes and both modes, checked against the oracle, then the select counters as one JSON line."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from datou_gomoku_muzero_b200.engine import SearchEngine  # noqa: E402
from oracle import oracle  # noqa: E402

out = {"cases": []}
tot = [0, 0, 0]
# (N, S, K, G, mode, logit_div (0 = dense / unquantised logits), accumulation dtype)
for N, S, K, G, mode, div, accum in [(15, 400, 16, 256, "AlphaZero", 16, "float64"), (15, 400, 16, 128, "MuZero", 16, "float64"),
                                     (9, 100, 16, 256, "AlphaZero", 4, "float64"), (19, 200, 32, 96, "AlphaZero", 2, "float64"),
                                     (6, 50, 8, 128, "MuZero", 16, "float64"), (15, 1200, 16, 32, "AlphaZero", 16, "float64"),
                                     (15, 400, 16, 256, "AlphaZero", 0, "float64"), (15, 400, 16, 256, "AlphaZero", 0, "float32"),
                                     (15, 400, 16, 128, "MuZero", 0, "float32"), (9, 100, 16, 256, "AlphaZero", 16, "float32")]:
    A = N * N
    rs = np.random.RandomState(N * 1000 + S)
    boards = np.zeros((G, A), np.int8); players = np.ones(G, np.int8)
    last = np.full(G, -1, np.int32); mc = np.zeros(G, np.int32)
    for g in range(G):
        k = int(rs.randint(0, A - 1)) if g % 7 else 0
        cells = rs.permutation(A)[:k]
        boards[g, cells[0::2]] = 1; boards[g, cells[1::2]] = -1
        players[g] = 1 if k % 2 == 0 else -1
        last[g] = cells[-1] if k else -1; mc[g] = k
    gum = rs.gumbel(0, 1, (G, A))
    eng = SearchEngine(G, board_size=N, num_simulations=S, num_top_actions=K, mode=mode, accum_dtype=accum)
    eng.set_roots(boards, players, last, mc)
    eng.search_e0(torch.from_numpy(gum).cuda(), 11, div)
    pol, val, act, vis = (t.cpu().numpy() for t in eng.finalize())
    cfg = oracle.make_config(board_size=N, num_simulations=S, num_top_actions=K, eval_seed=11, logit_div=div,
                             mode=0 if mode == "AlphaZero" else 1, accum_dtype=int(accum == "float32"))
    opol, oval, oact, ovis = oracle.search_batch(cfg, boards, players, last, mc, gum)
    fb, fast, bad = eng.select_counters()
    out["cases"].append(dict(N=N, S=S, K=K, G=G, mode=mode, logit_div=div, accum=accum, visits_equal=bool(np.array_equal(vis, ovis)),
                             moves_equal=bool(np.array_equal(act, oact)), values_equal=bool(np.array_equal(val, oval)),
                             fallback=fb, certified=fast, contradicted=bad))
    tot = [tot[0] + fb, tot[1] + fast, tot[2] + bad]
out["fallback"], out["certified"], out["contradicted"] = tot
print(json.dumps(out))
