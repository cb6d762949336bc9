"""Runs the UNMODIFIED reference (baseline/_ref, installed by baseline/install_reference.py) on the host
cores for bench.py's `--impl reference` arm and `cpu_baseline` leg.  Nothing here re-implements the
reference: its own AlphaZeroMCTS / MuZeroMCTS.search, universal_worker and inference_server_worker are
imported and called; the only additions are (a) the config overrides the benchmark's workload needs
(board size, simulations -- config.py ships 6x6 / MuZero), applied in each process before `game` is
imported (game.py binds its defaults at import), and (b) the fixed evaluator E0 behind the reference's own
queue protocol (tests/golden/e0_py.E0Queue, the same object that produced the golden vectors).

Measurements:
  tree_only()  -- P processes, each running reference `search(game)` with E0 in-process: the reference's tree
                  code with a free evaluator, the counterpart of the GPU engine's E0 numbers;
  topology()   -- the reference's production topology (main.py:91-104): `universal_worker` x W processes
                  + one `inference_server_worker` running GomokuNetEZ on the GPU, mp.Queue IPC per simulation;
  inprocess_net() -- BASELINE.md section 4.1 (E1): ONE reference process playing games with the reference's
                  GomokuNetEZ on the GPU answering its queue protocol synchronously in-process (no IPC);
  per()        -- reference InMemoryReplayBuffer.sample + update_priorities on one core.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")
STUBS = os.path.join(ROOT, "tests", "golden", "_stubs")      # empty seaborn / matplotlib (plotting is not on the path)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def available() -> bool:
    return os.path.exists(os.path.join(REF, "mcts.py"))


def _configure(N, n_in_row, S, K, mode):
    """Import the reference's config and point it at the benchmark workload (before `game` is imported)."""
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    for p in (GOLDEN, STUBS, REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    torch.set_num_threads(1)
    from config import config
    config.BOARD_SIZE, config.N_IN_ROW, config.ACTION_SPACE_SIZE = N, n_in_row, N * N
    config.NUM_SIMULATIONS, config.NUM_TOP_ACTIONS, config.MCTS_IMPLEMENTATION = S, K, mode
    return config


def _positions(n, N, seed):
    """The bench's synthetic mid-game roots (bench.staggered_positions), as reference GomokuGame objects."""
    import game as ref_game
    A = N * N
    rs = np.random.RandomState(seed)
    out = []
    for g in range(n):
        k = (g * 37) % 160 if N == 15 else (g * 7) % (A // 2)
        cells = rs.permutation(A)[:k]
        gm = ref_game.GomokuGame(board_size=N, n_in_row=5)
        gm.board.reshape(-1)[cells[0::2]] = 1
        gm.board.reshape(-1)[cells[1::2]] = -1
        gm.current_player = 1 if k % 2 == 0 else -1
        gm.last_move = (int(cells[-1]) // N, int(cells[-1]) % N) if k else None
        gm.move_count = k
        out.append(gm)
    return out


def _tree_worker(idx, n_workers, N, S, K, mode, e0_seed, logit_div, warmup, steps, start, done, out_q):
    try:
        _configure(N, 5, S, K, mode)
        import mcts as ref_mcts
        from e0_py import E0Queue
        q = E0Queue(seed=e0_seed, logit_div=logit_div)
        q.set_action_space(N * N)
        eng = (ref_mcts.AlphaZeroMCTS if mode == "AlphaZero" else ref_mcts.MuZeroMCTS)(idx, q, q)
        games = _positions(n_workers * 4, N, 1234)[idx::n_workers]
        np.random.seed(100 + idx)
        for i in range(warmup):
            eng.search(games[i % len(games)])
        start.wait()
        t0 = time.perf_counter()
        for i in range(steps):
            _, _, action = eng.search(games[i % len(games)])
            assert action >= 0
        dt = time.perf_counter() - t0
        done.wait()
        out_q.put((idx, dt, q.n_initial + q.n_recurrent))
    except Exception as ex:                      # never leave the parent waiting on a barrier
        out_q.put((idx, -1.0, repr(ex)))
        try:
            start.abort(); done.abort()
        except Exception:
            pass


def tree_only(n_workers, steps, warmup=1, N=15, S=400, K=16, mode="AlphaZero", e0_seed=2024, logit_div=16):
    """`n_workers` processes x `steps` reference searches each (one search per process per step)."""
    ctx = mp.get_context("spawn")
    start, done, out_q = ctx.Barrier(n_workers + 1), ctx.Barrier(n_workers + 1), ctx.Queue()
    procs = [ctx.Process(target=_tree_worker, args=(i, n_workers, N, S, K, mode, e0_seed, logit_div, warmup, steps, start, done, out_q),
                         daemon=True) for i in range(n_workers)]
    for p in procs:
        p.start()
    try:
        start.wait(timeout=600)
        t0 = time.perf_counter()
        done.wait(timeout=3600)
        dt = time.perf_counter() - t0
    except Exception:
        errs = []
        while not out_q.empty():
            errs.append(out_q.get())
        for p in procs:
            p.terminate()
        raise RuntimeError(f"reference workers failed: {errs}")
    res = [out_q.get(timeout=60) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    bad = [r for r in res if r[1] < 0]
    if bad:
        raise RuntimeError(f"reference workers failed: {bad}")
    searches = n_workers * steps
    return {"seconds": dt, "searches": searches, "sims_per_sec": searches * S / dt, "moves_per_sec": searches / dt,
            "processes": n_workers, "ms_per_step": dt / steps * 1e3,
            "per_process_sims_per_sec": float(np.mean([steps * S / r[1] for r in res]))}


# ---------------------------------------------------------------------------------------------------------
# the reference's production topology
# ---------------------------------------------------------------------------------------------------------
def _server_entry(N, S, K, mode, args):
    cfg = _configure(N, 5, S, K, mode)
    import torch
    torch.set_num_threads(max(1, min(4, (os.cpu_count() or 2) // 4)))
    cfg.DEVICE = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    import workers as ref_workers
    ref_workers.inference_server_worker(*args)


def _worker_entry(N, S, K, mode, args):
    _configure(N, 5, S, K, mode)
    import workers as ref_workers
    ref_workers.universal_worker(*args)


def topology(seconds=60.0, n_workers=None, N=15, S=400, K=16, mode="AlphaZero"):
    """main.py:91-104 without the trainer / data loader / UI: W `universal_worker` processes and one
    `inference_server_worker` (GomokuNetEZ on the GPU, random init published through model_update_queue exactly
    like the trainer does, workers.py:496-497).  Moves are counted from the SelfPlayMove messages each worker
    puts on ui_queue (workers.py:179)."""
    import tempfile
    cfg = _configure(N, 5, S, K, mode)
    import torch
    from ipc_messages import ModelWeightsUpdate, SelfPlayMove
    from network import GomokuNetEZ
    n_workers = n_workers or max(1, (os.cpu_count() or 4) - 2)
    ctx = mp.get_context("spawn")
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="gmz_ref_")
    os.chdir(tmp)                                            # the workers create outputs/ + SQLite files in the cwd
    procs = []
    try:
        request_queue, result_queues = ctx.Queue(), [ctx.Queue() for _ in range(n_workers)]
        model_update_queue, initial_req = ctx.Queue(maxsize=1), ctx.Queue()
        data_queue, log_status_queue, ui_queue = ctx.Queue(), ctx.Queue(), ctx.Queue(maxsize=100000)
        replay_q, trainer_ev_q, log_queue = ctx.Queue(), ctx.Queue(), ctx.Queue()
        shutdown, ready, pause = ctx.Event(), ctx.Event(), ctx.Event()
        worker_mode, latest = ctx.Value("i", 0), ctx.Value("i", 0)
        torch.manual_seed(0)
        net = GomokuNetEZ(cfg)
        model_update_queue.put(ModelWeightsUpdate({k: v.cpu() for k, v in net.state_dict().items()}))
        server = ctx.Process(target=_server_entry, args=(N, S, K, mode, (request_queue, result_queues, model_update_queue, initial_req,
                                                                         shutdown, ready, log_queue)), daemon=True)
        server.start(); procs.append(server)
        if not ready.wait(timeout=300):
            raise RuntimeError("reference inference server did not become ready")
        for i in range(n_workers):
            p = ctx.Process(target=_worker_entry, args=(N, S, K, mode, (i, worker_mode, data_queue, log_status_queue, ui_queue, shutdown,
                                                                        request_queue, result_queues[i], replay_q, trainer_ev_q, latest,
                                                                        log_queue, pause)), daemon=True)
            p.start(); procs.append(p)

        def drain():
            n = 0
            for q in (log_queue, log_status_queue, trainer_ev_q, replay_q, data_queue):
                try:
                    while True:
                        q.get_nowait()
                except Exception:
                    pass
            try:
                while True:
                    if isinstance(ui_queue.get_nowait(), SelfPlayMove):
                        n += 1
            except Exception:
                pass
            return n
        # steady state: wait for the first move of every worker generation, then count for `seconds`
        t_dead = time.perf_counter() + 300
        first = 0
        while first < 1 and time.perf_counter() < t_dead:
            time.sleep(0.25); first += drain()
        t0, moves = time.perf_counter(), 0
        while time.perf_counter() - t0 < seconds:
            time.sleep(0.25); moves += drain()
        dt = time.perf_counter() - t0
        alive = sum(p.is_alive() for p in procs)
        return {"seconds": dt, "moves": moves, "moves_per_sec": moves / dt, "sims_per_sec": moves * S / dt, "workers": n_workers,
                "processes_alive": alive, "device": str(torch.device("cuda" if torch.cuda.is_available() else "cpu")),
                "what": "reference universal_worker x %d + inference_server_worker (GomokuNetEZ 8x128 fp32 on the GPU), mp.Queue IPC "
                        "per simulation, %s mode %dx%d / %d sims" % (n_workers, mode, N, N, S)}
    finally:
        try:
            shutdown.set()
        except Exception:
            pass
        time.sleep(0.5)
        for p in procs:
            if p.is_alive():
                p.terminate()
        for p in procs:
            p.join(timeout=10)
        os.chdir(cwd)
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)


# ---------------------------------------------------------------------------------------------------------
# one reference process, network on the GPU answered in-process
# ---------------------------------------------------------------------------------------------------------
class _InProcessNet:
    """The evaluator protocol of mcts.py:66-82 answered synchronously by a model in this process (the role
    webui.py:107-141 plays for the UI): put() runs the network, get() hands the result back."""

    def __init__(self, model, device):
        self.model, self.device, self.pending, self.calls = model, device, None, 0

    def put(self, request):
        import torch
        _worker, kind, payload = request
        self.calls += 1
        with torch.no_grad():
            if kind == "initial":
                p, v, h = self.model.initial_inference(torch.from_numpy(payload)[None].to(self.device))
                self.pending = (p[0].cpu().numpy(), v[0, 0].cpu().numpy(), h.cpu().numpy())
            else:
                hidden, actions = payload
                out = self.model.recurrent_inference(torch.from_numpy(hidden).to(self.device),
                                                     torch.from_numpy(actions).long().to(self.device))
                self.pending = tuple(t.cpu().numpy() for t in out)

    def get(self, timeout=None):
        import queue
        if self.pending is None:
            raise queue.Empty
        out, self.pending = self.pending, None
        return out

    get_nowait = get


def _inprocess_entry(N, S, K, mode, seconds, out_q):
    try:
        cfg = _configure(N, 5, S, K, mode)
        import torch
        import game as ref_game
        import mcts as ref_mcts
        from network import GomokuNetEZ
        dev = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        cfg.DEVICE = dev
        torch.manual_seed(0)
        q = _InProcessNet(GomokuNetEZ(cfg).to(dev).eval(), dev)
        eng = (ref_mcts.AlphaZeroMCTS if mode == "AlphaZero" else ref_mcts.MuZeroMCTS)(0, q, q)
        np.random.seed(0)
        gm = ref_game.GomokuGame(board_size=N, n_in_row=5)
        eng.search(gm)                                        # warm-up: CUDA context, cuDNN plans
        moves = games = 0
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            _, _, action = eng.search(gm)
            assert action >= 0
            gm.do_move(action); moves += 1
            if gm.get_game_ended() is not None:
                gm.reset(); games += 1
        dt = time.perf_counter() - t0
        out_q.put({"seconds": dt, "moves": moves, "games_finished": games, "moves_per_sec": moves / dt, "sims_per_sec": moves * S / dt,
                   "evaluator_calls": q.calls, "device": str(dev),
                   "what": "ONE reference process: %sMCTS.search + GomokuGame loop, %dx%d / %d sims, reference GomokuNetEZ (fp32, "
                           "random init) on the %s answering the queue protocol in-process (no IPC)" % (mode, N, N, S, dev.type)})
    except Exception as ex:
        out_q.put({"error": repr(ex)})


def inprocess_net(seconds=10.0, N=9, S=100, K=16, mode="AlphaZero"):
    ctx = mp.get_context("spawn")
    out_q = ctx.Queue()
    p = ctx.Process(target=_inprocess_entry, args=(N, S, K, mode, seconds, out_q), daemon=True)
    p.start()
    try:
        res = out_q.get(timeout=seconds + 600)
    finally:
        p.join(timeout=30)
        if p.is_alive():
            p.terminate()
    if "error" in res:
        raise RuntimeError("reference in-process run failed: " + res["error"])
    return res


def per(capacity=1_000_000, batch=360, rounds=200, seed=0):
    """Reference InMemoryReplayBuffer (replay_buffer.py:43-106) at `capacity`, full, ENABLE_PER: `rounds` x
    (sample(batch) + update_priorities) on one core.  Returns samples/s."""
    for p in (REF,):
        if p not in sys.path:
            sys.path.insert(0, p)
    from config import config
    import replay_buffer as rb
    config.ENABLE_PER = True
    np.random.seed(seed)
    buf = rb.InMemoryReplayBuffer(capacity)
    # fill: leaves written directly and the inner nodes summed level by level (1 M reference add() calls take minutes
    # and are not what is timed); the tree that results is the one add() + update_priorities would have built
    pri = np.abs(np.random.randn(capacity)) + 1e-6
    t = buf.sum_tree
    t.tree[capacity - 1:] = pri
    for i in range(capacity - 2, -1, -1):
        t.tree[i] = t.tree[2 * i + 1] + t.tree[2 * i + 2]
    t.count, t.write_ptr = capacity, 0
    buf.data = np.zeros(capacity, np.int8)       # payloads are irrelevant to the tree arithmetic
    buf.sample(batch)
    t0 = time.perf_counter()
    for _ in range(rounds):
        _, idx, _w = buf.sample(batch)
        buf.update_priorities(idx, np.random.randn(batch).astype(np.float32))
    dt = time.perf_counter() - t0
    return {"seconds": dt, "samples_per_sec": rounds * batch / dt, "us_per_batch": dt / rounds * 1e6, "capacity": capacity, "batch": batch}
