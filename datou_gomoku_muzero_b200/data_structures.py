"""Trajectory contract types, field-for-field the reference's (data_structures.py:9-26):
the data-queue item is `(GameRecord, [TrainingSlice], model_version)` (workers.py:230)."""
from collections import namedtuple

# one finished game: per-move lists
GameRecord = namedtuple("GameRecord", "observations actions rewards policies values board_states")
# one training sample: obs f32 [U+1,3,N,N], actions i32 [U] (pad -1), rewards f32 [U],
# policies f64 [U+1,A] (pad 0), values f32 [U+1]
TrainingSlice = namedtuple("TrainingSlice", "observation action_history reward_history policy_history value_history")
