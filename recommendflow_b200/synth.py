"""Synthetic workloads of BASELINE.json / SURVEY.md §8(d), generated with numpy only.

C2 "embedding microbench": F hashed fields, keys "f{field:02d}_{v}" with v ~ U[0, 1e7)
(numpy default_rng(20260101 + field)), dense bags of L keys, tables U(-0.05, 0.05).
Keys are built straight into (uint8 arena, int32 offsets) -- no Python string objects.
"""
import numpy as np
import torch

from .strings import ARENA_SLACK, StringColumn


def decimal_keys(prefix: bytes, values: np.ndarray):
    """ASCII keys prefix + str(v) for non-negative ints -> (arena uint8, offsets int32[n+1])."""
    v = np.asarray(values, dtype=np.int64).ravel()
    nd = np.ones(v.size, dtype=np.int64)
    p = 10
    while True:
        more = v >= p
        if not more.any():
            break
        nd += more
        p *= 10
    plen = len(prefix)
    lens = nd + plen
    offs = np.zeros(v.size + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    arena = np.empty(int(offs[-1]), dtype=np.uint8)
    start = offs[:-1]
    for k, ch in enumerate(prefix):
        arena[start + k] = ch
    rest = v.copy()
    for k in range(int(nd.max())):
        sel = nd > k
        arena[start[sel] + plen + nd[sel] - 1 - k] = (rest[sel] % 10 + 48).astype(np.uint8)
        rest //= 10
    return arena, offs.astype(np.int32)


def c2_field_keys(field, batch, bag_len, seed_base=20260101, batch_index=0, zipf=None):
    rng = np.random.default_rng(seed_base + field + 1000 * batch_index)
    n = batch * bag_len
    if zipf:
        v = np.minimum(rng.zipf(zipf, size=n), 10**7) - 1
    else:
        v = rng.integers(0, 10**7, size=n)
    return decimal_keys(b"f%02d_" % field, v)


class PackedBatch(object):
    """All string fields of one batch packed into ONE arena and ONE offsets buffer, so a batch
    crosses PCIe as two copies.  `columns()` hands out per-field StringColumn views."""

    def __init__(self, data, offsets, layout):
        self.data, self.offsets, self.layout = data, offsets, layout   # layout: name -> (byte0, off0, n_items, shape)

    @staticmethod
    def pack(fields, pin=False):
        """fields: {name: (arena, offsets, shape)} (numpy)."""
        total_bytes = sum(int(a.size) for a, _, _ in fields.values())
        total_offs = sum(int(o.size) for _, o, _ in fields.values())
        data = torch.zeros(total_bytes + ARENA_SLACK, dtype=torch.uint8)
        offsets = torch.empty(total_offs, dtype=torch.int32)
        if pin:
            data, offsets = data.pin_memory(), offsets.pin_memory()
        layout, b0, o0 = {}, 0, 0
        for name, (arena, offs, shape) in fields.items():
            data[b0:b0 + arena.size] = torch.from_numpy(arena)
            offsets[o0:o0 + offs.size] = torch.from_numpy(offs)
            layout[name] = (b0, o0, int(offs.size) - 1, tuple(shape))
            b0 += int(arena.size)
            o0 += int(offs.size)
        return PackedBatch(data, offsets, layout)

    @property
    def nbytes(self):
        return int(self.data.numel()) + 4 * int(self.offsets.numel())

    def to(self, device, non_blocking=True, out=None):
        """Copy to the device (into `out`, a PackedBatch of device buffers of the same size, if given)."""
        if out is None:
            return PackedBatch(self.data.to(device, non_blocking=non_blocking),
                               self.offsets.to(device, non_blocking=non_blocking), self.layout)
        out.data.copy_(self.data, non_blocking=non_blocking)
        out.offsets.copy_(self.offsets, non_blocking=non_blocking)
        out.layout = self.layout
        return out

    def columns(self):
        cols = {}
        for name, (b0, o0, n, shape) in self.layout.items():
            cols[name] = StringColumn(self.data[b0:], self.offsets[o0:o0 + n + 1], shape)
        return cols
