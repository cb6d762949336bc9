import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from bench import staggered_positions, N, S, K_TOP, A, E0_SEED, LOGIT_DIV
from datou_gomoku_muzero_b200.engine import SearchEngine
G=4096
eng = SearchEngine(G, board_size=N, num_simulations=S, num_top_actions=K_TOP)
eng.set_roots(*staggered_positions(G, 0))
gum = torch.empty((G, A), dtype=torch.float64, device="cuda"); eng.fill_gumbel(gum, 1, 0)
ta, td = eng.search_e0(gum, E0_SEED, LOGIT_DIV, trace=True)
d = td[:, :S-1].float()
tot = d.sum(1).cpu().numpy()
print("per-game sum of leaf depths: mean %.0f  min %.0f  p50 %.0f  p90 %.0f p99 %.0f max %.0f  max/mean %.2f" % (tot.mean(), tot.min(), np.percentile(tot,50), np.percentile(tot,90), np.percentile(tot,99), tot.max(), tot.max()/tot.mean()))
print("max depth overall", int(d.max().item()), "mean depth", float(d.mean().item()))
mc = staggered_positions(G,0)[3]
for lo,hi in [(0,20),(20,60),(60,100),(100,160)]:
    m=(mc>=lo)&(mc<hi); print(f"stones {lo}-{hi}: mean total depth {tot[m].mean():.0f}")
