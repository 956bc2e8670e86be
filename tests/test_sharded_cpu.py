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
