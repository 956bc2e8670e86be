"""Shared helpers for the GPU parity tests: seeded synthetic keys + oracle-side mirrors."""
import numpy as np
import torch

import oracle
from recommendflow_b200.strings import StringColumn


def random_strings(rng, n, max_len=24, alphabet=None, empty_frac=0.0):
    """n random byte strings; lengths uniform in [0, max_len]; a fraction forced empty (pads)."""
    lens = rng.integers(0, max_len + 1, size=n)
    if empty_frac:
        lens[rng.uniform(size=n) < empty_frac] = 0
    offs = np.zeros(n + 1, dtype=np.int64)
    offs[1:] = np.cumsum(lens)
    if alphabet is None:
        arena = rng.integers(0, 256, size=int(offs[-1]), dtype=np.uint8)
    else:
        arena = np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), size=int(offs[-1]))]
    return arena.astype(np.uint8), offs.astype(np.int32)


def column(arena, offs, shape, bag_offsets=None, device="cuda"):
    return StringColumn.from_arena(arena, offs, shape, bag_offsets).to(device)


def tables(rng, n_tables, rows, dim):
    return [rng.uniform(-0.05, 0.05, size=(rows, dim)).astype(np.float32) for _ in range(n_tables)]


def to_dev(ws):
    return [torch.from_numpy(w).cuda() for w in ws]
