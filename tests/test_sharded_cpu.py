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


# ---- C5: sharded tables + data-parallel towers + all-gathered in-batch softmax, under gloo -------------------------
def _c5_setup(seed=5):
    from recommendflow_b200.backend.blocks.mlp import create_mlp
    torch.manual_seed(seed)
    user_tower = create_mlp([16, 8], 0.0, "selu", None, name="user_tower")
    ad_tower = create_mlp([16, 8], 0.0, "selu", None, name="ad_tower")
    x = torch.zeros(2, 8)
    user_tower(x), ad_tower(x)                     # builds the (seeded) dense variables on the CPU
    return user_tower, ad_tower


def _c5_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from recommendflow_b200.training_sharded import ShardedRecallTrainer
        from tests.shard_util import TorchLossOps
        N, D, B = 211, 8, 24
        full = {n: np.random.default_rng(i).uniform(-0.05, 0.05, size=(N, D)).astype(np.float32) for i, n in enumerate(("u", "a"))}
        bags = {}
        for n in ("u", "a"):
            bags[n] = ShardedEmbeddingBag(N, D, combiner="avg", salt=None, mask_value="", transport="nccl", max_batch=B,
                                          max_keys=B * 6, device="cpu", ops=OracleShardOps())
            bags[n].set_full_weights(full[n])
        user_tower, ad_tower = _c5_setup()
        trainer = ShardedRecallTrainer({"u": bags["u"]}, {"a": bags["a"]}, user_tower, ad_tower, learning_rate=1e-2,
                                       loss_ops=TorchLossOps())
        losses = []
        for step in (1, 2, 3):
            batch = {}
            for i, n in enumerate(("u", "a")):
                arena, offs, bag = rank_batch(rank, B, 6, seed=100 * step + 7 * i)
                batch[n] = StringColumn.from_arena(arena, offs, (B, None), bag)
            losses.append(float(trainer.train_step(batch, np.ones(B, np.float32))))
        dense = [p.detach().numpy().copy() for p in trainer.dense_opt.params]
        q.put((rank, losses, dense, {n: b.shard.detach().numpy().copy() for n, b in bags.items()}))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_c5_train_step_matches_single_process():
    """3 steps of the C5 trainer on 2 ranks == 3 steps of ONE process holding the full tables and the global batch
    (same towers, Keras Adam everywhere, in-batch softmax over all 2B docs): losses, dense variables and table shards
    agree to fp32 re-association (the sharded forward sums partial pools per owner, gradients are all-reduced)."""
    world, N, D, B = 2, 211, 8, 24
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_c5_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {r: (l, d, s) for r, l, d, s in (q.get(timeout=180) for _ in procs)}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # ---- single process: full tables, global batch, torch autograd + the oracle's sparse Keras Adam ----
    from recommendflow_b200.training import KerasAdam
    full = {n: np.random.default_rng(i).uniform(-0.05, 0.05, size=(N, D)).astype(np.float32) for i, n in enumerate(("u", "a"))}
    mom = {n: (np.zeros_like(full[n]), np.zeros_like(full[n])) for n in full}
    user_tower, ad_tower = _c5_setup()
    params = []
    for t in (user_tower, ad_tower):
        for p in t.parameters():
            p.requires_grad_(True)
            params.append(p)
    opt = KerasAdam(params, learning_rate=1e-2)
    want_losses = []
    for step in (1, 2, 3):
        leaves, ids_all, bag_all = {}, {}, {}
        for i, n in enumerate(("u", "a")):
            pooled, ids_l, offs_l = [], [], [0]
            for rank in range(world):
                arena, offs, bag = rank_batch(rank, B, 6, seed=100 * step + 7 * i)
                ids = oracle.hash_strings(arena, offs, N, "", None)
                pooled.append(sharded_reference(ids, bag, full[n], world, "avg"))
                ids_l.append(ids)
                offs_l += (bag[1:].astype(np.int64) + offs_l[-1]).tolist()
            leaves[n] = torch.from_numpy(np.concatenate(pooled)).requires_grad_(True)
            ids_all[n], bag_all[n] = np.concatenate(ids_l), np.asarray(offs_l, dtype=np.int32)
        u = torch.nn.functional.normalize(user_tower(leaves["u"]), dim=1, eps=1e-12)
        a = torch.nn.functional.normalize(ad_tower(leaves["a"]), dim=1, eps=1e-12)
        s = 20.0 * (u @ a.t())
        loss = torch.mean(-(torch.diagonal(s) - torch.logsumexp(s, dim=1)))
        opt.zero_grad()
        loss.backward()
        want_losses.append(float(loss))
        opt.step()
        for n in ("u", "a"):
            cnt = np.maximum(np.diff(bag_all[n]), 1).astype(np.float32)[:, None]
            oracle.bag_backward_adam(ids_all[n], leaves[n].grad.numpy() / cnt, full[n], mom[n][0], mom[n][1], step, lr=1e-2,
                                     combiner="sum", bag_offsets=bag_all[n])
    for rank in range(world):
        losses, dense, shards = res[rank]
        np.testing.assert_allclose(losses, want_losses, rtol=2e-5, atol=1e-6)
        for got, p in zip(dense, params):
            np.testing.assert_allclose(got, p.detach().numpy(), rtol=1e-4, atol=2e-6)
        for n in ("u", "a"):
            # Adam normalises the gradient (m / (sqrt(v) + eps)): where a row's gradient is ~1e-8, fp32 re-association of
            # the gradient moves the update by a visible fraction of the step (lr = 1e-2): allow 0.5 % of one full step
            np.testing.assert_allclose(shards[n], full[n][rank::world], rtol=1e-4, atol=5e-5)
    assert want_losses[-1] < want_losses[0]
