"""One short network-evaluated search (G = 4096, 15x15, 8x128 bf16) for per-kernel timing under
`ncu --metrics gpu__time_duration.sum`: python tools/netstep_probe.py [sims] [graph 0|1]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import staggered_positions, N, S, K_TOP, A, N_IN_ROW
from datou_gomoku_muzero_b200.config import Config
from datou_gomoku_muzero_b200.engine import SearchEngine
from datou_gomoku_muzero_b200.network import GomokuNetEZ, NetworkSearch

sims = int(sys.argv[1]) if len(sys.argv) > 1 else 12
graph = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
G = 4096
torch.manual_seed(0)
eng = SearchEngine(G, board_size=N, n_in_row=N_IN_ROW, num_simulations=S, num_top_actions=K_TOP, accum_dtype="float32")
cfg = Config(BOARD_SIZE=N, ACTION_SPACE_SIZE=A, NUM_RES_BLOCKS=8, NUM_FILTERS=128, HEAD_HIDDEN_DIM=64)
ns = NetworkSearch(eng, GomokuNetEZ(cfg), dtype=torch.bfloat16, graph=graph)
gum = torch.empty((G, A), dtype=torch.float64, device="cuda")
eng.fill_gumbel(gum, 1, 0)
eng.set_roots(*staggered_positions(G, 0))
ns.search(gum, num_simulations=6)
eng.set_roots(*staggered_positions(G, 0))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ns.search(gum, num_simulations=sims); e1.record(); torch.cuda.synchronize()
print(f"{sims} sims graph={graph}: {e0.elapsed_time(e1) / sims:.3f} ms per simulation step")
if os.environ.get("GMZ_TORCH_PROFILE"):
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        ns.search(gum, num_simulations=11)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
