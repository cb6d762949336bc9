import sys; sys.path.insert(0,'/root/repo')
import torch
from datou_gomoku_muzero_b200.engine import SearchEngine
from datou_gomoku_muzero_b200.selfplay import SelfPlayEngine
from datou_gomoku_muzero_b200.trajectory import TrajectoryStore
G=4096
eng = SearchEngine(G, board_size=15, num_simulations=64)
sp = SelfPlayEngine(eng, "e0", seed=3, noise_seed=4)
traj = TrajectoryStore(eng, extra_slots=512)
off = eng._ws_ptr - eng.workspace.data_ptr()
gs = eng.workspace[off:off + G*1024].view(torch.int32).reshape(G, 256)
for i in range(4):
    sp.play(moves_per_game=6, traj=traj); torch.cuda.synchronize()
    busy = gs[:, 908//4]; parked = gs[:, 912//4]; winner = gs[:, 904//4]; active = gs[:, 884//4]; mc = gs[:, 880//4]
    fin = traj.harvest(copy_policies=False)
    m,f = eng.play_counters()

    print("launch",i,"moves",m,"finished",f,"unserved",eng.tickets_unserved, "busy",int(busy.sum()),"parked",int(parked.sum()),
          "winner!=2",int((winner!=2).sum()),"active",int(active.sum()), "move_count min/max", int(mc.min()), int(mc.max()), flush=True)
