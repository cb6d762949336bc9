"""Development probe: run single bench legs in isolation (is a leg's number sensitive to what ran before it?).
    python tools/leg_probe.py muzero | selfplay | net"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

which = sys.argv[1] if len(sys.argv) > 1 else "muzero"
dev = torch.device("cuda:0")
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
if which == "muzero":
    for i in range(2):
        r = bench.muzero_leg(dev, peaks, 4096)
        print("muzero", i, r["ms_per_search"], r["sims_per_sec"], r["tensor"]["frac"], r["e0_fused"]["sims_per_sec"], flush=True)
        torch.cuda.empty_cache()
elif which == "net":
    from datou_gomoku_muzero_b200.engine import SearchEngine
    torch.backends.cudnn.benchmark = os.environ.get("GMZ_CUDNN_BENCHMARK", "0") != "0"
    for name in sys.argv[2:] or ["bf16"]:
        eng = SearchEngine(4096, board_size=bench.N, n_in_row=bench.N_IN_ROW, num_simulations=bench.S, num_top_actions=bench.K_TOP,
                           device=dev, accum_dtype="float32")
        r = bench.net_leg(eng, dev, peaks, name)
        print("net", name, "benchmark", torch.backends.cudnn.benchmark, r["sims_per_sec"], r["net_forward_ms"], r["clocks"], r["tensor"]["frac_sustained"], flush=True)
        del eng
elif which == "selfplay":
    from datou_gomoku_muzero_b200.engine import SearchEngine
    from datou_gomoku_muzero_b200.replay_buffer import DeviceReplayBuffer
    from datou_gomoku_muzero_b200.selfplay import SelfPlayEngine
    from datou_gomoku_muzero_b200.trajectory import TrajectoryStore
    G = 4096
    eng = SearchEngine(G, board_size=bench.N, num_simulations=bench.S, device=dev)
    sp = SelfPlayEngine(eng, "e0", seed=bench.E0_SEED, noise_seed=2000)
    eng.set_roots(*bench.staggered_positions(G, 0))
    traj = TrajectoryStore(eng, extra_slots=G // 2)
    buf = DeviceReplayBuffer(400_000, bench.N, device=dev)
    def t(fn):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return r, (time.perf_counter() - t0) * 1e3
    for rep in range(4):
        _, a = t(lambda: eng.selfplay_e0(G * 24, bench.E0_SEED, bench.LOGIT_DIV, 2000, traj, True))
        pg, b = t(lambda: traj.pack_finished(recycle=True))
        _, c = t(lambda: buf.add_packed(pg))
        (batch, idx, w), d = t(lambda: buf.sample(360))
        _, e = t(lambda: buf.update_priorities(idx, batch[4][:, 0] - 0.5))
        print(f"rep {rep}: play {a:.2f} ms, pack {b:.2f} ms ({len(pg)} games, {pg.n_moves} records), add_packed {c:.2f}, sample {d:.2f}, update {e:.2f}", flush=True)
    # pack_finished, piece by piece
    import ctypes as C, numpy as np
    from datou_gomoku_muzero_b200.config import config
    from datou_gomoku_muzero_b200._lib import check
    for rep in range(3):
        eng.selfplay_e0(G * 24, bench.E0_SEED, bench.LOGIT_DIV, 2000, traj, True)
        torch.cuda.synchronize()
        tt = [time.perf_counter()]
        def mark(): torch.cuda.synchronize(); tt.append(time.perf_counter())
        n = int(traj.fin_count.item()); mark()
        q = traj.fin_queue[:n].cpu().numpy(); mark()
        lens = np.minimum(q[:, 2], traj.max_moves).astype(np.int64); off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64); mark()
        rec = torch.empty((int(off[-1]), buf.stride), dtype=torch.uint8, device=dev); mark()
        off_d = torch.as_tensor(off[:-1].copy(), device=dev)
        dpow = torch.tensor([config.DISCOUNT ** i for i in range(config.N_STEPS + 1)], dtype=torch.float64, device=dev); mark()
        check(eng.lib.gmz_traj_pack(C.byref(traj.c), eng.N, traj.fin_queue.data_ptr(), n, off_d.data_ptr(), dpow.data_ptr(),
                                    int(config.N_STEPS), rec.data_ptr(), eng._stream()), "gmz_traj_pack"); mark()
        traj.fin_count.zero_(); traj.release(torch.as_tensor(q[:, 0].astype(np.int32), device=dev)); mark()
        d = [(b - a) * 1e3 for a, b in zip(tt, tt[1:])]
        print("pack pieces ms: count %.3f, table d2h %.3f, host offsets %.3f, alloc %.3f (%.0f MB), h2d %.3f, kernel %.3f, release %.3f" %
              (d[0], d[1], d[2], d[3], rec.numel() / 1e6, d[4], d[5], d[6]), flush=True)
        del rec
    for rep in range(2):
        r = bench.selfplay_e2e_leg(dev, 0, 1, G, 24)
        print("leg", rep, r["seconds"], r["moves_per_sec"], flush=True)
