"""Kernel micro-bench for development: time the fused search kernel (and optionally the stepwise
path) on the bench workload and spot-check parity against the oracle on the first games.
    GMZ_LIB=/path/to/variant.so python tools/kbench.py [--games 4096] [--iters 5] [--check 64]"""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import staggered_positions, N, S, K_TOP, A, E0_SEED, LOGIT_DIV
from datou_gomoku_muzero_b200.engine import SearchEngine

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=4096)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--check", type=int, default=64)
ap.add_argument("--stepwise", action="store_true")
ap.add_argument("--mode", default="AlphaZero", choices=["AlphaZero", "MuZero"])
ap.add_argument("--selfplay", type=int, default=0, help="time the persistent self-play kernel for this many moves per game")
args = ap.parse_args()
G = args.games
eng = SearchEngine(G, board_size=N, num_simulations=S, num_top_actions=K_TOP, mode=args.mode)
roots = staggered_positions(G, 0)
eng.set_roots(*roots)
gum = torch.empty((G, A), dtype=torch.float64, device="cuda")
eng.fill_gumbel(gum, 1, 0)
run = (lambda: eng.search_stepwise_e0(gum, E0_SEED, LOGIT_DIV)) if args.stepwise else (lambda: eng.search_e0(gum, E0_SEED, LOGIT_DIV))
for _ in range(2):
    run()
torch.cuda.synchronize()
if args.selfplay:
    from datou_gomoku_muzero_b200.trajectory import TrajectoryStore
    traj = TrajectoryStore(eng, extra_slots=G // 2)
    eng.selfplay_e0(G, E0_SEED, LOGIT_DIV, 3, traj, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    m0, _ = eng.play_counters()
    e0.record(); eng.selfplay_e0(G * args.selfplay, E0_SEED, LOGIT_DIV, 3, traj, True); e1.record(); torch.cuda.synchronize()
    m1, f1 = eng.play_counters()
    ms = e0.elapsed_time(e1)
    print(f"lib={os.environ.get('GMZ_LIB','default')} G={G} selfplay {args.selfplay} moves/game: {ms:.2f} ms  "
          f"{(m1-m0)*S/ms/1e3:.2f} M sims/s  {(m1-m0)/ms*1e3:.0f} moves/s  moves {m1-m0} finished {f1}")
    sys.exit(0)
ts = []
for _ in range(args.iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts))
print(f"lib={os.environ.get('GMZ_LIB','default')} G={G} {'stepwise' if args.stepwise else 'fused'}: {ms:.3f} ms/search  "
      f"{G*S/ms/1e3:.2f} M sims/s  (min {min(ts):.3f} max {max(ts):.3f})")
if args.check:
    from oracle import oracle
    pol, val, act, vis = (t.cpu().numpy() for t in eng.finalize())
    c = args.check
    cfg = oracle.make_config(board_size=N, num_simulations=S, num_top_actions=K_TOP, eval_seed=E0_SEED, logit_div=LOGIT_DIV,
                             mode=0 if args.mode == "AlphaZero" else 1)
    g = gum.cpu().numpy()
    opol, oval, oact, ovis = oracle.search_batch(cfg, roots[0][:c], roots[1][:c], roots[2][:c], roots[3][:c], g[:c])
    ok = np.array_equal(vis[:c], ovis) and np.array_equal(act[:c], oact)
    print(f"parity on first {c} games: visits/actions {'bit-exact' if ok else 'MISMATCH'}; value bit-exact {np.array_equal(val[:c], oval)}; "
          f"max |dpolicy| {np.abs(pol[:c]-opol).max():.2e}")
    if not ok:
        bad = [i for i in range(c) if not np.array_equal(vis[i], ovis[i])]
        print("mismatching games:", bad[:10]); sys.exit(1)
