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


def _train_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from recommendflow_b200.sharded import ShardedEmbeddingBag
        from recommendflow_b200.strings import StringColumn
        N, D, B = 5003, 32, 256
        full = np.random.default_rng(1).uniform(-0.05, 0.05, size=(N, D)).astype(np.float32)
        layer = ShardedEmbeddingBag(N, D, combiner="avg", salt=None, mask_value="", transport="nccl", max_batch=B,
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


def test_sharded_backward_adam_multi_gpu():
    # same construction as tests/test_sharded_cpu.py::test_two_rank_gloo_sharded_backward_adam, on the CUDA kernels
    world, N, D, B = 2, 5003, 32, 256
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, world, port, q)) for r in range(world)]
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
