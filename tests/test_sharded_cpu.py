"""world_size-2 gloo run of the row-sharded forward's host logic (exchange choreography, buffer
layout, split sizes) with oracle-backed compute ops standing in for the CUDA kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from recommendflow_b200.sharded import ShardedEmbeddingBag, shard_rows
from recommendflow_b200.strings import StringColumn
from tests.shard_util import OracleShardOps, rank_batch, sharded_reference


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, combiner, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N, D, B = 1009, 8, 64
        full = np.random.default_rng(1).uniform(-0.05, 0.05, size=(N, D)).astype(np.float32)
        layer = ShardedEmbeddingBag(N, D, combiner=combiner, salt=[2022, 2023], mask_value="", transport="nccl",
                                    max_batch=B, max_keys=B * 20, device="cpu", ops=OracleShardOps())
        layer.set_full_weights(full)
        assert layer.shard.shape[0] == shard_rows(N, rank, world)
        arena, offs, bag = rank_batch(rank, B, 20)
        col = StringColumn.from_arena(arena, offs, (B, None), bag)
        got = layer(col).numpy()
        ids = oracle.hash_strings(arena, offs, N, "", [2022, 2023])
        want = sharded_reference(ids, bag, full, world, combiner)
        seq = oracle.bag_pool(ids, full, combiner, bag_offsets=bag)
        q.put((rank, bool(np.array_equal(got, want)), float(np.abs(got - seq).max())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("combiner", ["sum", "avg", "max"])
def test_two_rank_gloo_sharded_forward(combiner):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, combiner, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, exact, err in res:
        assert exact, f"rank {rank}: sharded result differs from the sharded restatement"
        # vs the single-GPU sequential sum: only fp32 re-association (<= 20 adds of |x| <= 0.05)
        assert err <= 20 * 0.05 * 2.0 ** -22, (rank, err)


def test_shard_rows_partition_the_table():
    for N in (1, 7, 8, 1009, 100_000_000):
        for W in (1, 2, 8):
            assert sum(shard_rows(N, r, W) for r in range(W)) == N


def _train_worker(rank, world, port, combiner, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N, D, B = 211, 8, 48
        full = np.random.default_rng(1).uniform(-0.05, 0.05, size=(N, D)).astype(np.float32)
        layer = ShardedEmbeddingBag(N, D, combiner=combiner, salt=None, mask_value="", transport="nccl",
                                    max_batch=B, max_keys=B * 12, device="cpu", ops=OracleShardOps())
        layer.set_full_weights(full)
        for step in (1, 2):
            arena, offs, bag = rank_batch(rank, B, 12, seed=100 * step)
            layer(StringColumn.from_arena(arena, offs, (B, None), bag))
            g = np.random.default_rng(7 * step + rank).standard_normal((B, D)).astype(np.float32)
            layer.apply_adam(torch.from_numpy(g), learning_rate=1e-2)
        q.put((rank, layer.shard.detach().numpy().copy(), layer._adam["m"].numpy().copy(), layer._adam["v"].numpy().copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("combiner", ["sum", "avg"])
def test_two_rank_gloo_sharded_backward_adam(combiner):
    # two optimisation steps on a row-sharded table == the same two Keras-Adam steps on the full table with the
    # keys of rank 0 then rank 1 (per row that is exactly the order every owner sums in), bit for bit
    world, N, D, B = 2, 211, 8, 48
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, world, port, combiner, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {r: (w, m, v) for r, w, m, v in (q.get(timeout=120) for _ in procs)}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = np.random.default_rng(1).uniform(-0.05, 0.05, size=(N, D)).astype(np.float32)
    m, v = np.zeros_like(full), np.zeros_like(full)
    for step in (1, 2):
        ids_all, grads_all, offs_all = [], [], [0]
        for rank in range(world):
            arena, offs, bag = rank_batch(rank, B, 12, seed=100 * step)
            ids = oracle.hash_strings(arena, offs, N, "", None)
            g = np.random.default_rng(7 * step + rank).standard_normal((B, D)).astype(np.float32)
            if combiner == "avg":
                g = g / np.maximum(np.diff(bag), 1).astype(np.float32)[:, None]
            ids_all.append(ids)
            grads_all.append(g)
            base = offs_all[-1]
            offs_all += (bag[1:].astype(np.int64) + base).tolist()
        oracle.bag_backward_adam(np.concatenate(ids_all), np.concatenate(grads_all), full, m, v, step, lr=1e-2,
                                 combiner="sum", bag_offsets=np.asarray(offs_all, dtype=np.int32))
    for rank in range(world):
        w_r, m_r, v_r = res[rank]
        assert np.array_equal(w_r.view(np.uint32), full[rank::world].view(np.uint32)), rank
        assert np.array_equal(m_r.view(np.uint32), m[rank::world].view(np.uint32)), rank
        assert np.array_equal(v_r.view(np.uint32), v[rank::world].view(np.uint32)), rank
