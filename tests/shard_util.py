"""Helpers for the sharded-path tests: a numpy restatement of the sharded algorithm, and
oracle-backed compute ops so the exchange choreography can run on CPU under gloo."""
import ctypes

import numpy as np
import torch

import oracle
from recommendflow_b200.strings import StringColumn


def sharded_reference(ids, bag_offsets, full_table, world, combiner):
    """Numpy restatement of the sharded forward for ONE source rank: partial pools per owner in key
    order (fp32), combined in rank order; avg divides by the bag's key count."""
    B = len(bag_offsets) - 1
    D = full_table.shape[1]
    out = np.zeros((B, D), dtype=np.float32)
    for b in range(B):
        keys = ids[bag_offsets[b]:bag_offsets[b + 1]]
        if len(keys) == 0:
            continue
        parts = []
        for g in range(world):
            mine = keys[keys % world == g]
            if combiner in ("sum", "avg"):
                acc = np.zeros(D, dtype=np.float32)
                for k in mine:
                    acc = acc + full_table[k]
            else:
                acc = np.full(D, np.inf if combiner == "min" else -np.inf, dtype=np.float32)
                for k in mine:
                    acc = np.minimum(acc, full_table[k]) if combiner == "min" else np.maximum(acc, full_table[k])
            parts.append(acc)
        acc = parts[0]
        for g in range(1, world):
            if combiner in ("sum", "avg"):
                acc = acc + parts[g]
            else:
                acc = np.minimum(acc, parts[g]) if combiner == "min" else np.maximum(acc, parts[g])
        if combiner == "avg":
            acc = acc / np.float32(len(keys))
        out[b] = acc
    return out


def _write(ptr, arr):
    arr = np.ascontiguousarray(arr)
    ctypes.memmove(ptr, arr.ctypes.data, arr.nbytes)


class OracleShardOps(object):
    """CPU stand-ins for CudaShardOps (tests only): same contracts, numpy + oracle arithmetic."""

    def hash(self, keys, num_bins, mask_value, salt):
        assert isinstance(keys, StringColumn)
        ids = oracle.hash_strings(keys.data.numpy()[:keys.nbytes], keys.offsets.numpy(), num_bins, mask_value, salt)
        return torch.from_numpy(ids)

    def route(self, ids, bag_offsets, bag_len, batch, world, counts_ws, offs_local, offs_dst_ptrs, rows_dst_ptrs):
        ids = ids.numpy()
        bo = bag_offsets.numpy() if bag_offsets is not None else np.arange(batch + 1, dtype=np.int64) * bag_len
        bag_of = np.repeat(np.arange(batch), np.diff(bo))
        for g in range(world):
            sel = ids % world == g
            counts = np.bincount(bag_of[sel], minlength=batch).astype(np.int32)
            offs = np.zeros(batch + 1, dtype=np.int32)
            offs[1:] = np.cumsum(counts)
            _write(offs_dst_ptrs[g], offs)
            _write(rows_dst_ptrs[g], (ids[sel] // world).astype(np.int64))     # stable: key order kept

    def pool(self, shard, rows_per_src, offs_per_src, outs_per_src, batch, combiner, est_items):
        w = shard.numpy()
        for rows, offs, out in zip(rows_per_src, offs_per_src, outs_per_src):
            offs = offs.numpy()
            res = oracle.bag_pool(rows.numpy()[:offs[-1]], w, combiner, bag_offsets=offs)
            if combiner in ("min", "max"):
                res[np.diff(offs) == 0] = np.inf if combiner == "min" else -np.inf
            out.copy_(torch.from_numpy(res))

    def combine(self, partials, world, batch, dim, combiner, bag_len, bag_offsets, out):
        p = partials.numpy()
        acc = p[0].copy()
        for g in range(1, world):
            acc = acc + p[g] if combiner in ("sum", "avg") else (np.minimum(acc, p[g]) if combiner == "min" else np.maximum(acc, p[g]))
        cnt = np.diff(bag_offsets.numpy()) if bag_offsets is not None else np.full(batch, bag_len)
        if combiner == "avg":
            acc = acc / np.maximum(cnt, 1).astype(np.float32)[:, None]
        acc[cnt == 0] = 0
        out.copy_(torch.from_numpy(acc.astype(np.float32)))

    def adam(self, shard, m, v, rows, offs_all, grads, state, params):
        oracle.bag_backward_adam(rows.numpy(), grads.numpy(), shard.numpy(), m.numpy(), v.numpy(), params["step"],
                                 lr=params["learning_rate"], beta1=params["beta_1"], beta2=params["beta_2"], eps=params["epsilon"],
                                 combiner="sum", bag_offsets=offs_all.numpy(), lazy=params["lazy"])


def rank_batch(rank, B, max_len, seed=4242, alphabet=b"abcdefghijklmnopqrstuvwxyz0123456789_"):
    """Jagged keys of one rank: lengths ~ U{0..max_len} (seed 4242 + rank)."""
    rng = np.random.default_rng(seed + rank)
    lens = rng.integers(0, max_len + 1, size=B)
    bag = np.zeros(B + 1, dtype=np.int32)
    bag[1:] = np.cumsum(lens)
    n = int(bag[-1])
    klen = rng.integers(1, 13, size=n)
    offs = np.zeros(n + 1, dtype=np.int32)
    offs[1:] = np.cumsum(klen)
    arena = np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), size=int(offs[-1]))]
    return arena.copy(), offs, bag


class TorchLossOps(object):
    """CPU stand-ins for training_sharded.CudaLossOps (tests only): the [B, B] block primitives in plain torch."""

    def block_lse(self, q, a, scale):
        return torch.logsumexp(scale * (q @ a.t()), dim=1)

    def rowdot(self, q, a):
        return (q * a).sum(dim=1)

    def block_grads(self, q, a, y, lse, scale, upstream, own_block):
        B = q.shape[0]
        c = torch.exp(scale * (q @ a.t()) - lse[:, None])
        if own_block:
            c = c - torch.eye(B, dtype=q.dtype)
        c = c * (upstream * scale / B) * y[:, None]
        return c @ a, c.t() @ q
