"""Trajectories of the persistent self-play kernel -> the reference's data-queue items.

`TrajectoryStore` owns the device buffers the kernel records into (include/gmz.h `gmz_traj`) and
harvests finished games; `build_game_record` / `cut_training_slices` turn one finished game into
the `(GameRecord, [TrainingSlice])` pair the reference's self-play loop emits (workers.py:183-230),
with the same dtypes and the same arithmetic (final-reward pattern, n-step value targets, padding).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import GmzTraj, check
from .config import config
from .data_structures import GameRecord, TrainingSlice


class TrajectoryStore:
    def __init__(self, engine, extra_slots=None, max_moves=None):
        e = self.e = engine
        G, A = e.G, e.A
        self.n_slots = G + (max(64, G // 4) if extra_slots is None else int(extra_slots))
        self.max_moves = A if max_moves is None else int(max_moves)
        dev = e.device
        n, T = self.n_slots, self.max_moves
        self.policy = torch.zeros((n, T, A), dtype=torch.float64, device=dev)
        self.value = torch.zeros((n, T), dtype=torch.float64, device=dev)
        self.action = torch.full((n, T), -1, dtype=torch.int32, device=dev)
        self.start_board = torch.zeros((n, 2, 8), dtype=torch.int64, device=dev)
        self.start_info = torch.zeros((n, 4), dtype=torch.int32, device=dev)
        self.free_slots = torch.zeros(n, dtype=torch.int32, device=dev)
        self.free_top = torch.zeros(1, dtype=torch.int32, device=dev)
        self.fin_queue = torch.zeros((n, 4), dtype=torch.int32, device=dev)
        self.fin_count = torch.zeros(1, dtype=torch.int32, device=dev)
        p = lambda t: t.data_ptr()
        self.c = GmzTraj(n, T, n, 0, p(self.policy), p(self.value), p(self.action), p(self.start_board),
                         p(self.start_info), p(self.free_slots), p(self.free_top), p(self.fin_queue), p(self.fin_count))
        check(e.lib.gmz_traj_init(e.handle, C.byref(self.c), e._stream()), "gmz_traj_init")
        e.launches += 1

    def harvest(self, copy_policies=True):
        """Finished games since the last harvest -> list of dicts (host arrays); recycles their slots
        and restarts parked games."""
        e = self.e
        n = int(self.fin_count.item())
        if n == 0:
            return []
        q = self.fin_queue[:n].cpu().numpy()
        slots = torch.as_tensor(q[:, 0].astype(np.int64), device=e.device)
        Tmax = int(q[:, 2].max())
        act = self.action[slots, :Tmax].cpu().numpy()
        val = self.value[slots, :Tmax].cpu().numpy()
        pol = self.policy[slots, :Tmax].cpu().numpy() if copy_policies else None
        sb = self.start_board[slots].cpu().numpy().view(np.uint64)
        si = self.start_info[slots].cpu().numpy()
        out = []
        for i in range(n):
            T = int(q[i, 2])
            bits = np.unpackbits(sb[i].view(np.uint8).reshape(2, 64), axis=1, bitorder="little")[:, :e.A]
            board = bits[0].astype(np.int8) - bits[1].astype(np.int8)
            out.append(dict(game=int(q[i, 1]), length=T, winner=int(q[i, 3]), actions=act[i, :T].copy(),
                            values=val[i, :T].copy(), policies=None if pol is None else pol[i, :T].copy(),
                            start_board=board.reshape(e.N, e.N), start_player=int(si[i, 0]),
                            start_move_count=int(si[i, 1]), start_last_move=int(si[i, 2])))
        # recycle: push the harvested slots back on the free stack, clear the queue, restart parked games
        top = int(self.free_top.item())
        self.free_slots[top:top + n] = slots.to(torch.int32)
        self.free_top.fill_(top + n)
        self.fin_count.zero_()
        check(e.lib.gmz_selfplay_unpark(e.handle, C.byref(self.c), e._stream()), "gmz_selfplay_unpark")
        e.launches += 1
        return out


def final_rewards(num_moves, winner):
    """workers.py:183-187: last mover +1, then -1, then r[i] = -r[i+2] backwards; zeros on a draw."""
    r = np.zeros(num_moves, dtype=np.float32)
    if winner != 0 and num_moves > 0:
        r[-1] = 1.0
        if num_moves > 1:
            r[-2] = -1.0
        for i in range(num_moves - 3, -1, -1):
            r[i] = -r[i + 2]
    return r


def compute_n_step_returns(rewards, values, discount, n_steps):
    """workers.py:144-152, same operand types: `rewards` a list of Python floats, values -> float32
    array (so the bootstrap product is float32), Python `sum` for the reward part."""
    T = len(rewards)
    returns = np.zeros(T, dtype=np.float32)
    v32 = np.array(values, dtype=np.float32)
    powers = [discount ** i for i in range(n_steps)]
    gamma_n = discount ** n_steps
    for t in range(T - 1, -1, -1):
        b = t + n_steps
        boot = v32[b] * gamma_n if b < len(v32) else 0.0
        acc = sum(powers[i] * rewards[t + i] for i in range(n_steps) if t + i < T)
        returns[t] = acc + boot
    return returns.tolist()


def build_game_record(rec, board_size=None):
    """One harvested game (from an empty start) -> GameRecord (workers.py:172-206)."""
    N = board_size or rec["start_board"].shape[0]
    T = rec["length"]
    board = rec["start_board"].astype(np.int8).copy()
    player, last = rec["start_player"], rec["start_last_move"]
    observations, board_states = [], []
    for t in range(T):
        obs = np.zeros((3, N, N), dtype=np.float32)
        obs[0][board == player] = 1.0
        obs[1][board == -player] = 1.0
        if last >= 0:
            obs[2, last // N, last % N] = 1.0
        observations.append(obs)
        board_states.append(board.copy())
        a = int(rec["actions"][t])
        board[a // N, a % N] = player
        player, last = -player, a
    rewards = final_rewards(T, rec["winner"]).tolist()
    values = [np.float64(v) for v in rec["values"]]
    targets = compute_n_step_returns(rewards, values, config.DISCOUNT, config.N_STEPS)
    policies = [rec["policies"][t] for t in range(T)]
    return GameRecord(observations, [int(a) for a in rec["actions"]], rewards, policies, targets, board_states)


def cut_training_slices(game_record):
    """workers.py:208-222: one slice per move, U unroll steps, padded with zeros / -1."""
    U = config.NUM_UNROLL_STEPS
    obs, acts, rews, pols, vals = (game_record.observations, game_record.actions, game_record.rewards,
                                   game_record.policies, game_record.values)
    T = len(acts)
    if T == 0:
        return []
    obs_p = np.concatenate([np.stack(obs), np.zeros((U + 1,) + obs[0].shape, obs[0].dtype)])
    act_p = np.array(list(acts) + [-1] * U, dtype=np.int32)
    rew_p = np.array(list(rews) + [0.0] * U, dtype=np.float32)
    pol_p = np.concatenate([np.stack(pols), np.zeros((U + 1,) + pols[0].shape, pols[0].dtype)])
    val_p = np.array(list(vals) + [0.0] * (U + 1), dtype=np.float32)
    return [TrainingSlice(obs_p[i:i + U + 1].copy(), act_p[i:i + U].copy(), rew_p[i:i + U].copy(),
                          pol_p[i:i + U + 1].copy(), val_p[i:i + U + 1].copy()) for i in range(T)]
