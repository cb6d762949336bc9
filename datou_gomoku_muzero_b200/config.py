"""Runtime configuration singleton, attribute-compatible with the reference's `config`
(reference config.py:4-109): modules read attributes at call time and tests mutate them,
so this is a plain mutable object, not a frozen dataclass.  Only the attributes the search /
self-play / replay path reads are listed; values are the reference's shipped defaults.
"""
from __future__ import annotations

_DEFAULTS = dict(
    # game + search (config.py:18-34)
    BOARD_SIZE=6, N_IN_ROW=5, NUM_SIMULATIONS=400, NUM_TOP_ACTIONS=16, MCTS_IMPLEMENTATION="MuZero",
    C_VISIT=30, C_SCALE=1.0, VALUE_MINMAX_DELTA=1e-3, DISCOUNT=0.997,
    # network shape (config.py:39-51), used by network.GomokuNetEZ
    VALUE_SUPPORT_MIN=-1, VALUE_SUPPORT_MAX=1, VALUE_SUPPORT_BINS=3,
    REWARD_SUPPORT_MIN=-1, REWARD_SUPPORT_MAX=1, REWARD_SUPPORT_BINS=3,
    NUM_RES_BLOCKS=8, NUM_FILTERS=128, HEAD_HIDDEN_DIM=64,
    # trajectory / replay contracts (config.py:56-59, 71, 95-102)
    PHYSICAL_BATCH_SIZE=360, TRAIN_BUFFER_SIZE=1000000, NUM_UNROLL_STEPS=5, N_STEPS=10,
    ENABLE_PER=False, PER_ALPHA=0.6, PER_BETA=0.4, PER_BETA_INCREMENT=0.00001, PER_EPSILON=1e-6,
    INFERENCE_BATCH_SIZE=15, NUM_WORKERS=15, MODEL_UPDATE_INTERVAL=1000,
)


class Config:
    def __init__(self, **overrides):
        for k, v in _DEFAULTS.items():
            setattr(self, k, v)
        self.ACTION_SPACE_SIZE = self.BOARD_SIZE * self.BOARD_SIZE
        for k, v in overrides.items():
            setattr(self, k, v)

    @property
    def DEVICE(self):
        import torch
        return torch.device("cuda" if torch.cuda.is_available() else "cpu")


config = Config()
