#!/usr/bin/env python
"""Diagnostics for the sharded path on N GPUs: per-rank phase times, pooling with local vs peer
partial buffers, and raw exchange bandwidth (NCCL all_to_all, peer copies)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from recommendflow_b200.sharded import ShardedEmbeddingBag
    from recommendflow_b200.strings import StringColumn
    from tools.bench_sharded import jagged_keys

    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B, N, D = int(os.environ.get("DIAG_BATCH", "65536")), 100_000_000, 128
    arena, offs, bag = jagged_keys(rank, B, 200)
    col = StringColumn.from_arena(arena, offs, (B, None), bag).to(dev)
    layer = ShardedEmbeddingBag(N, D, combiner="avg", transport="p2p", max_batch=B, max_keys=col.n_items + 100000)
    out = torch.empty(B, D, device=dev)
    for _ in range(3):
        layer(col, out=out)
    torch.cuda.synchronize()
    dist.barrier()

    def phases(tag):
        acc = {}
        for _ in range(5):
            layer.profile = []
            layer(col, out=out)
            torch.cuda.synchronize()
            for (n0, e0), (n1, e1) in zip(layer.profile[:-1], layer.profile[1:]):
                acc[n1] = acc.get(n1, 0.0) + e0.elapsed_time(e1) / 5
        layer.profile = None
        keys = ["route", "barrier0", "pool", "barrier1", "combine"]
        t = torch.tensor([acc[k] for k in keys], device=dev)
        allt = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        if rank == 0:
            print(tag, json.dumps({k: [round(float(a[i]), 3) for a in allt] for i, k in enumerate(keys)}))

    phases("peer partials  :")
    # variant: owners write partials into their OWN buffer (no NVLink stores in the pool kernel)
    b = layer._bufs
    saved = b["peer_partials"]
    b["peer_partials"] = [[b["partials"][j] for _ in range(world)] for j in range(layer.N_SETS)]
    phases("local partials :")
    b["peer_partials"] = saved

    # raw exchange bandwidth
    send = torch.randn(world, B, D, device=dev)
    recv = torch.empty_like(send)
    for _ in range(3):
        dist.all_to_all_single(recv, send)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        dist.all_to_all_single(recv, send)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    vol = (world - 1) * B * D * 4
    if rank == 0:
        print(f"nccl all_to_all_single {send.numel() * 4 / 1e6:.0f} MB: {ms:.3f} ms -> {vol / ms / 1e6:.0f} GB/s out per rank")
    # peer copies through the symmetric buffers (cudaMemcpy peer, one per destination)
    hdl = b["hdl"]
    peers = [hdl.get_buffer(r, (world, B, D), torch.float32, 0) for r in range(world)]
    dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        for r in range(world):
            if r != rank:
                peers[r][rank].copy_(send[r], non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    if rank == 0:
        print(f"peer copy_ x{world - 1} ({B * D * 4 / 1e6:.0f} MB each): {ms:.3f} ms -> {vol / ms / 1e6:.0f} GB/s out per rank")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
