"""Parity soak: many random searches on the GPU vs the CPU oracle, all games compared.
    python tools/soak_parity.py [--batches 25] [--games 4096]
Reports the number of searches whose visit counts / move / value differ (expected: 0)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from datou_gomoku_muzero_b200.engine import SearchEngine
from oracle import oracle

ap = argparse.ArgumentParser()
ap.add_argument("--batches", type=int, default=25)
ap.add_argument("--games", type=int, default=4096)
ap.add_argument("--board", type=int, default=15)
ap.add_argument("--sims", type=int, default=400)
ap.add_argument("--top", type=int, default=16)
ap.add_argument("--mode", default="AlphaZero", choices=["AlphaZero", "MuZero"])
ap.add_argument("--divs", default="16,16,4,2", help="logit_div values to draw from (0 = dense / unquantised logits and values)")
ap.add_argument("--accum", default="float64", choices=["float64", "float32"], help="tree arithmetic dtype (float32 = production)")
args = ap.parse_args()
DIVS = [int(x) for x in args.divs.split(",")]
N, S, K, G = args.board, args.sims, args.top, args.games
A = N * N
eng = SearchEngine(G, board_size=N, num_simulations=S, num_top_actions=K, mode=args.mode, accum_dtype=args.accum)
rs = np.random.RandomState(2026)
tot = bad_vis = bad_act = bad_val = 0
worst_pol = 0.0
t0 = time.time()
for b in range(args.batches):
    seed, div = int(rs.randint(1 << 30)), int(rs.choice(DIVS))
    boards = np.zeros((G, A), np.int8); players = np.ones(G, np.int8)
    last = np.full(G, -1, np.int32); mc = np.zeros(G, np.int32)
    ks = rs.randint(0, A, size=G)
    for g in range(G):
        cells = rs.permutation(A)[:ks[g]]
        boards[g, cells[0::2]] = 1; boards[g, cells[1::2]] = -1
        players[g] = 1 if ks[g] % 2 == 0 else -1
        last[g] = cells[-1] if ks[g] else -1; mc[g] = ks[g]
    gum = rs.gumbel(0, 1, (G, A))
    eng.set_roots(boards, players, last, mc)
    eng.search_e0(torch.from_numpy(gum).cuda(), seed, div)
    pol, val, act, vis = (t.cpu().numpy() for t in eng.finalize())
    cfg = oracle.make_config(board_size=N, num_simulations=S, num_top_actions=K, eval_seed=seed, logit_div=div,
                             mode=0 if args.mode == "AlphaZero" else 1, accum_dtype=int(args.accum == "float32"))
    opol, oval, oact, ovis = oracle.search_batch(cfg, boards, players, last, mc, gum)
    tot += G
    bad_vis += int((vis != ovis).any(axis=1).sum()); bad_act += int((act != oact).sum()); bad_val += int((val != oval).sum())
    worst_pol = max(worst_pol, float(np.abs(pol - opol).max()))
    print(f"batch {b}: seed {seed} logit_div {div}: cumulative {tot} searches, visit mismatches {bad_vis}, move {bad_act}, value {bad_val}", flush=True)
print(json.dumps({"searches": tot, "simulations": tot * S, "config": f"{N}x{N}, {S} sims, K={K}, {args.mode} mode, {args.accum} accumulation, random positions 0..{A - 1} stones, logit_div in {sorted(set(DIVS))} (0 = dense)",
                  "visit_count_mismatch_rate": bad_vis / tot, "move_mismatch_rate": bad_act / tot,
                  "visit_count_mismatches": bad_vis, "move_mismatches": bad_act, "value_bit_mismatches": bad_val,
                  "max_abs_policy_diff": worst_pol, "seconds": time.time() - t0,
                  "select_counters": dict(zip(("fallback_to_exact", "certified", "certified_but_exact_differs"), eng.select_counters()))}))
