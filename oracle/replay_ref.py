"""Restatement of TransReplayBuffer (utils/replay_buffer.py:3-30) for rows of named fields.

TEST INFRASTRUCTURE (see oracle/__init__.py).  A Python list FIFO exactly like the reference:
add_experience pops the oldest entry when full (:23-27), get_batch returns `batch_size`
CONSECUTIVE entries from a uniformly drawn start (:14-21), clear empties it (:29-30).
PARITY: pinned -- tests/golden/ref_replay.npz records the reference class itself (imported from
utils/replay_buffer.py by tests/golden/make_ref_golden.py) under a sequence of adds and seeded
get_batch calls; tests/test_oracle_predictor.py replays it here, tests/test_gpu_predictor.py on the
device ring.
"""
import numpy as np


class RefTransReplayBuffer:
    def __init__(self, size):
        self.size = size
        self.buffer = []

    def add_experience(self, trans):
        if 1 + len(self.buffer) > self.size:
            self.buffer.pop(0)
        self.buffer.append(trans)

    def add_rows(self, fields):
        """n transitions at once: fields = {name: array[n, width]}."""
        n = next(iter(fields.values())).shape[0]
        for i in range(n):
            self.add_experience({k: np.asarray(v[i], dtype=np.float32).reshape(-1) for k, v in fields.items()})

    def get_batch(self, batch_size, rng=np.random):
        sample_range = len(self.buffer) - batch_size + 1
        start = rng.choice(sample_range, 1, replace=False)[0]
        return [self.buffer[i + start] for i in range(batch_size)], int(start)

    def clear(self):
        self.buffer = []

    def __len__(self):
        return len(self.buffer)
