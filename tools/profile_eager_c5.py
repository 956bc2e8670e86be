#!/usr/bin/env python
"""cProfile of the eager C5 training step on rank 0 (torchrun, >= 2 GPUs): host time of the sharded trainer.
    torchrun --nproc-per-node 2 tools/profile_eager_c5.py [p2p|nccl]"""
import cProfile
import io
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    import torch
    import torch.distributed as dist
    from bench_sharded import jagged_keys
    from recommendflow_b200.backend.blocks.mlp import create_mlp
    from recommendflow_b200.sharded import ShardedEmbeddingBag
    from recommendflow_b200.strings import StringColumn
    from recommendflow_b200.training_sharded import ShardedRecallTrainer
    transport = sys.argv[1] if len(sys.argv) > 1 else "p2p"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    n_feat, D, B, max_len, N = 4, 16, 8192, 20, 20_000_000
    names = [f"user_{i}" for i in range(2)] + [f"ad_{i}" for i in range(2)]
    bags = {n: ShardedEmbeddingBag(N, D, combiner="avg", salt=None, mask_value="", transport=transport, max_batch=B, max_keys=B * max_len)
            for n in names}
    torch.manual_seed(11)
    towers = [create_mlp([256, 128], 0.0, "selu", None, name=t) for t in ("user_tower", "ad_tower")]
    x = torch.zeros(2, D * 2)
    for t in towers:
        t(x)
        t.to(dev)
    trainer = ShardedRecallTrainer({n: bags[n] for n in names[:2]}, {n: bags[n] for n in names[2:]}, towers[0], towers[1], learning_rate=1e-3)
    batches = []
    for bi in range(2):
        b = {}
        for i, n in enumerate(names):
            arena, offs, bag = jagged_keys(rank + 100 * i, B, max_len, bi)
            b[n] = StringColumn.from_arena(arena, offs, (B, None), bag).to(dev)
        batches.append(b)
    y = torch.ones(B, device=dev)
    for i in range(4):
        trainer.train_step(batches[i % 2], y)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10):
        trainer.train_step(batches[i % 2], y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    pr = cProfile.Profile()
    pr.enable()
    for i in range(20):
        trainer.train_step(batches[i % 2], y)
    pr.disable()
    torch.cuda.synchronize()
    if rank == 0:
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(40)
        print(f"transport {transport}: {ms:.3f} ms/step (eager, CUDA events)")
        print(s.getvalue())
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
