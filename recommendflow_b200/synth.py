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


# ---- closed-form table: any row can be rebuilt anywhere (host or device), bit for bit -----------------------------
# w[r, c] = float32(h) * float32(0.1 / 2^24) - float32(0.05),  h = ((r * 2654435761 + c * 40503 + 12345) mod 2^32) >> 8
# h < 2^24 converts to fp32 exactly; the multiply and the subtract are two separate IEEE fp32 operations on both
# sides (no FMA contraction: separate torch kernels / separate numpy ufuncs), so the device table equals the host
# rows bit for bit.  Used by the sharded benchmark's parity check: a rank checks its first bags against the oracle
# without anyone ever holding the 51 GB table on the host.
_CF_SCALE = np.float32(0.1 / 2 ** 24)
_CF_SHIFT = np.float32(0.05)


def closed_form_rows(row_ids, dim):
    """numpy: the [len(row_ids), dim] fp32 rows of the closed-form table."""
    r = np.asarray(row_ids, dtype=np.uint64).reshape(-1, 1)
    c = np.arange(dim, dtype=np.uint64).reshape(1, -1)
    h = ((r * np.uint64(2654435761) + c * np.uint64(40503) + np.uint64(12345)) & np.uint64(0xffffffff)) >> np.uint64(8)
    w = h.astype(np.float32)
    w *= _CF_SCALE
    w -= _CF_SHIFT
    return w


def fill_closed_form(table, first_row=0, row_stride=1, chunk_rows=1 << 20):
    """torch (any device): table[i] = closed-form row (first_row + i * row_stride); in place, chunked."""
    n, dim = table.shape
    c = torch.arange(dim, dtype=torch.int64, device=table.device).view(1, -1) * 40503 + 12345
    for i0 in range(0, n, chunk_rows):
        i1 = min(n, i0 + chunk_rows)
        r = (torch.arange(i0, i1, dtype=torch.int64, device=table.device) * row_stride + first_row).view(-1, 1)
        h = ((r * 2654435761 + c) & 0xffffffff) >> 8
        w = h.to(torch.float32)
        w.mul_(float(_CF_SCALE))
        w.sub_(float(_CF_SHIFT))
        table[i0:i1].copy_(w)
    return table
