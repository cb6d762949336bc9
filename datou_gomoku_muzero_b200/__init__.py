"""B200-native batched Gumbel-MCTS self-play engine for Gomoku (AlphaZero / MuZero modes).

Drop-in for the search hot path of Datou/Datou-gomoku-muzero: same class interfaces
(`AlphaZeroMCTS`, `MuZeroMCTS`, `GomokuGame`, `InMemoryReplayBuffer`, `GameRecord`,
`TrainingSlice`), trees held in HBM and advanced by hand-written sm_100a kernels behind the
C ABI in include/gmz.h.  Importing the package does not touch CUDA; constructing an engine does.
"""
from .config import Config, config  # noqa: F401
from .data_structures import GameRecord, TrainingSlice  # noqa: F401
from .game import GomokuGame  # noqa: F401

__all__ = ["config", "Config", "GomokuGame", "GameRecord", "TrainingSlice"]
