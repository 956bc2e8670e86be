"""Row-sharded forward on >= 2 real GPUs (NCCL + NVLink peer memory), both transports, vs the
numpy restatement of the sharded algorithm (bit-exact) and vs the sequential oracle (tolerance).
Skipped on a single-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_sharded_gpu.py -m gpu`."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from tests.shard_util import rank_batch, sharded_reference

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from recommendflow_b200.sharded import ShardedEmbeddingBag
        from recommendflow_b200.strings import StringColumn
        N, D, B = 100003, 128, 512
        full = np.random.default_rng(1).uniform(-0.05, 0.05, size=(N, D)).astype(np.float32)
        arena, offs, bag = rank_batch(rank, B, 200)
        col = StringColumn.from_arena(arena, offs, (B, None), bag).to(f"cuda:{rank}")
        ids = oracle.hash_strings(arena, offs, N, "", None)
        out = {}
        for combiner in ("avg", "sum", "max"):
            want = sharded_reference(ids, bag, full, world, combiner)
            seq = oracle.bag_pool(ids, full, combiner, bag_offsets=bag)
            for transport in ("p2p", "nccl"):
                layer = ShardedEmbeddingBag(N, D, combiner=combiner, salt=None, mask_value="", transport=transport,
                                            max_batch=B, max_keys=B * 200)
                layer.set_full_weights(full)
                for _ in range(3):                       # repeated steps reuse the exchange buffers
                    got = layer(col)
                torch.cuda.synchronize()
                got = got.cpu().numpy()
                out[(combiner, transport)] = (bool(np.array_equal(got, want)), float(np.abs(got - seq).max()))
        # combine-free mode: the owners reduce into the source's buffer (red.global.add over NVLink); summation order
        # across owners is free, so the check is the re-association bound against the sequential oracle
        from recommendflow_b200.sharded import BucketIds
        for combiner in ("avg", "sum"):
            seq = oracle.bag_pool(ids, full, combiner, bag_offsets=bag)
            layer = ShardedEmbeddingBag(N, D, combiner=combiner, salt=None, mask_value="", max_batch=B, max_keys=B * 200,
                                        deterministic=False)
            layer.set_full_weights(full)
            for _ in range(3):
                got = layer(col)
            # pipelined prepare / finish on the three streams gives the same values
            t0 = layer.prepare(col)
            t1 = layer.prepare(col)
            a = layer.finish(t0).clone()
            b = layer.finish(t1).clone()
            torch.cuda.synchronize()
            err = max(float(np.abs(x.cpu().numpy() - seq).max()) for x in (got, a, b))
            out[(combiner, "p2p-accumulate")] = (True, err)
        # pre-hashed ids (SURVEY.md §8d C4 "pre-hashed path"): routing skips the hash
        want = sharded_reference(ids, bag, full, world, "avg")
        layer = ShardedEmbeddingBag(N, D, combiner="avg", salt=None, mask_value="", max_batch=B, max_keys=B * 200)
        layer.set_full_weights(full)
        got = layer(BucketIds(torch.from_numpy(ids).to(f"cuda:{rank}"), col.bag_offsets)).cpu().numpy()
        out[("avg", "p2p-prehashed")] = (bool(np.array_equal(got, want)), 0.0)
        # the whole step through the single C-ABI entry (rf_sharded_bag_forward: own barriers, no torch collective)
        from recommendflow_b200.sharded import CAbiShardedStep, shard_rows
        for combiner in ("avg", "max"):
            stepper = CAbiShardedStep(N, D, combiner, B, B * 200)
            mine = torch.from_numpy(np.ascontiguousarray(full[rank::world])).to(f"cuda:{rank}")
            assert mine.shape[0] == shard_rows(N, rank, world)
            for _ in range(3):
                got = stepper(col, mine)
            torch.cuda.synchronize()
            want = sharded_reference(ids, bag, full, world, combiner)
            out[(combiner, "c-abi")] = (bool(np.array_equal(got.cpu().numpy(), want)), 0.0)
            got = stepper(BucketIds(torch.from_numpy(ids).to(f"cuda:{rank}"), col.bag_offsets), mine)
            torch.cuda.synchronize()
            out[(combiner, "c-abi-prehashed")] = (bool(np.array_equal(got.cpu().numpy(), want)), 0.0)
        # dense [B, L] padded input (reference semantics: pads pool row 0 of owner 0)
        L = 6
        a2, o2 = oracle.encode_strings([f"k{rank}_{i % 37}" if i % 5 else "" for i in range(B * L)])
        col2 = StringColumn.from_arena(a2, o2, (B, L)).to(f"cuda:{rank}")
        layer = ShardedEmbeddingBag(N, D, combiner="avg", salt=[2022, 2022], mask_value="", max_batch=B, max_keys=B * L)
        layer.set_full_weights(full)
        got = layer(col2).cpu().numpy()
        ids2 = oracle.hash_strings(a2, o2, N, "", [2022, 2022])
        want2 = sharded_reference(ids2, np.arange(B + 1) * L, full, world, "avg")
        out[("dense", "p2p")] = (bool(np.array_equal(got, want2)), 0.0)
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_sharded_forward_multi_gpu(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out in res:
        for key, (exact, err) in out.items():
            assert exact, f"rank {rank} {key}: differs from the sharded restatement"
            assert err <= 200 * 0.05 * 2.0 ** -21, (rank, key, err)     # fp32 re-association of <= 200 adds


def _train_worker(rank, world, port, q, transport="nccl"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from recommendflow_b200.sharded import ShardedEmbeddingBag
        from recommendflow_b200.strings import StringColumn
        N, D, B = 5003, 32, 256
        full = np.random.default_rng(1).uniform(-0.05, 0.05, size=(N, D)).astype(np.float32)
        layer = ShardedEmbeddingBag(N, D, combiner="avg", salt=None, mask_value="", transport=transport, max_batch=B,
                                    max_keys=B * 30)
        layer.set_full_weights(full)
        for step in (1, 2):
            arena, offs, bag = rank_batch(rank, B, 30, seed=100 * step)
            layer(StringColumn.from_arena(arena, offs, (B, None), bag).to(f"cuda:{rank}"))
            g = np.random.default_rng(7 * step + rank).standard_normal((B, D)).astype(np.float32)
            layer.apply_adam(torch.from_numpy(g).cuda(), learning_rate=1e-2)
        torch.cuda.synchronize()
        q.put((rank, layer.shard.detach().cpu().numpy(), layer._adam["m"].cpu().numpy(), layer._adam["v"].cpu().numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["nccl", "p2p"])
def test_sharded_backward_adam_multi_gpu(transport):
    # same construction as tests/test_sharded_cpu.py::test_two_rank_gloo_sharded_backward_adam, on the CUDA kernels; p2p: the
    # gradients reach the owners by peer stores and the routed rows are compacted out of the forward's (gapped) exchange set
    world, N, D, B = 2, 5003, 32, 256
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, world, port, q, transport)) for r in range(world)]
    for p in procs:
        p.start()
    res = {r: (w, m, v) for r, w, m, v in (q.get(timeout=300) for _ in procs)}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = np.random.default_rng(1).uniform(-0.05, 0.05, size=(N, D)).astype(np.float32)
    m, v = np.zeros_like(full), np.zeros_like(full)
    for step in (1, 2):
        ids_all, grads_all, offs_all = [], [], [0]
        for rank in range(world):
            arena, offs, bag = rank_batch(rank, B, 30, seed=100 * step)
            g = np.random.default_rng(7 * step + rank).standard_normal((B, D)).astype(np.float32)
            g = g / np.maximum(np.diff(bag), 1).astype(np.float32)[:, None]
            ids_all.append(oracle.hash_strings(arena, offs, N, "", None))
            grads_all.append(g)
            base = offs_all[-1]
            offs_all += (bag[1:].astype(np.int64) + base).tolist()
        assert np.bincount(np.concatenate(ids_all), minlength=N).max() <= 128       # every run is summed in key order
        oracle.bag_backward_adam(np.concatenate(ids_all), np.concatenate(grads_all), full, m, v, step, lr=1e-2,
                                 combiner="sum", bag_offsets=np.asarray(offs_all, dtype=np.int32))
    for rank in range(world):
        for got, want in zip(res[rank], (full, m, v)):
            assert np.array_equal(got.view(np.uint32), want[rank::world].view(np.uint32)), rank


# ---- C5: sharded tables + data-parallel towers + all-gathered in-batch softmax on 2 GPUs ------------------------------
def _c5_gpu_worker(rank, world, port, q, transport="nccl"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import recommendflow_b200.dense_ops as dense_ops
        from recommendflow_b200.backend.blocks.mlp import create_mlp
        from recommendflow_b200.sharded import ShardedEmbeddingBag
        from recommendflow_b200.strings import StringColumn
        from recommendflow_b200.training_sharded import ShardedRecallTrainer
        dense_ops.DEFAULT_PRECISION = "fp32"              # exact-fp32 loss kernels: the comparison below is tight
        N, D, B = 2003, 16, 256
        full = {n: np.random.default_rng(i).uniform(-0.05, 0.05, size=(N, D)).astype(np.float32) for i, n in enumerate(("u", "a"))}
        bags = {}
        for n in ("u", "a"):
            bags[n] = ShardedEmbeddingBag(N, D, combiner="avg", salt=None, mask_value="", transport=transport, max_batch=B, max_keys=B * 8)
            bags[n].set_full_weights(full[n])
        torch.manual_seed(5)
        towers = [create_mlp([32, 16], 0.0, "selu", None, name=t) for t in ("user_tower", "ad_tower")]
        x = torch.zeros(2, D)
        for t in towers:
            t(x)                                          # seeded CPU init, identical on every rank ...
            t.to(dev)                                     # ... then moved
        trainer = ShardedRecallTrainer({"u": bags["u"]}, {"a": bags["a"]}, towers[0], towers[1], learning_rate=1e-2)
        losses = []
        for step in (1, 2, 3):
            batch = {}
            for i, n in enumerate(("u", "a")):
                arena, offs, bag = rank_batch(rank, B, 8, seed=100 * step + 7 * i)
                batch[n] = StringColumn.from_arena(arena, offs, (B, None), bag).to(dev)
            losses.append(float(trainer.train_step(batch, torch.ones(B, device=dev))))
        torch.cuda.synchronize()
        q.put((rank, losses, [p.detach().cpu().numpy() for p in trainer.dense_opt.params],
               {n: b.shard.detach().cpu().numpy() for n, b in bags.items()}))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["nccl", "p2p"])
def test_c5_train_step_multi_gpu_matches_single_process(transport):
    """tests/test_sharded_cpu.py::test_two_rank_gloo_c5_train_step_matches_single_process on the CUDA kernels and NCCL:
    3 steps on 2 GPUs vs ONE CPU process with the full tables and the global batch."""
    from recommendflow_b200.backend.blocks.mlp import create_mlp
    from recommendflow_b200.training import KerasAdam
    world, N, D, B = 2, 2003, 16, 256
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_c5_gpu_worker, args=(r, world, port, q, transport)) for r in range(world)]
    for p in procs:
        p.start()
    res = {r: (l, d, s) for r, l, d, s in (q.get(timeout=300) for _ in procs)}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = {n: np.random.default_rng(i).uniform(-0.05, 0.05, size=(N, D)).astype(np.float32) for i, n in enumerate(("u", "a"))}
    mom = {n: (np.zeros_like(full[n]), np.zeros_like(full[n])) for n in full}
    torch.manual_seed(5)
    towers = [create_mlp([32, 16], 0.0, "selu", None, name=t) for t in ("user_tower", "ad_tower")]
    x = torch.zeros(2, D)
    params = []
    for t in towers:
        t(x)
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
                arena, offs, bag = rank_batch(rank, B, 8, seed=100 * step + 7 * i)
                ids = oracle.hash_strings(arena, offs, N, "", None)
                pooled.append(sharded_reference(ids, bag, full[n], world, "avg"))
                ids_l.append(ids)
                offs_l += (bag[1:].astype(np.int64) + offs_l[-1]).tolist()
            leaves[n] = torch.from_numpy(np.concatenate(pooled)).requires_grad_(True)
            ids_all[n], bag_all[n] = np.concatenate(ids_l), np.asarray(offs_l, dtype=np.int32)
        u = torch.nn.functional.normalize(towers[0](leaves["u"]), dim=1, eps=1e-12)
        a = torch.nn.functional.normalize(towers[1](leaves["a"]), dim=1, eps=1e-12)
        s = 20.0 * (u @ a.t())
        loss = torch.mean(-(torch.diagonal(s) - torch.logsumexp(s, dim=1)))
        opt.zero_grad()
        loss.backward()
        want_losses.append(float(loss.detach()))
        opt.step()
        for n in ("u", "a"):
            cnt = np.maximum(np.diff(bag_all[n]), 1).astype(np.float32)[:, None]
            oracle.bag_backward_adam(ids_all[n], leaves[n].grad.numpy() / cnt, full[n], mom[n][0], mom[n][1], step, lr=1e-2,
                                     combiner="sum", bag_offsets=bag_all[n])
    for rank in range(world):
        losses, dense, shards = res[rank]
        np.testing.assert_allclose(losses, want_losses, rtol=1e-4, atol=1e-5)
        # Adam normalises gradients: tiny-gradient elements move by a visible fraction of a step under fp32 reordering
        for got, p in zip(dense, params):
            np.testing.assert_allclose(got, p.detach().numpy(), rtol=2e-3, atol=2e-4)
        for n in ("u", "a"):
            np.testing.assert_allclose(shards[n], full[n][rank::world], rtol=2e-3, atol=2e-4)
    assert want_losses[-1] < want_losses[0]
