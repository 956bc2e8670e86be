#!/usr/bin/env python
"""C4: jagged behaviour sequences (1..max_len keys), mean pooling, one
rows x dim fp32 table row-sharded (id % world) across the GPUs of one box.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_sharded.py [...]
    python tools/bench_sharded.py            # N = 1: the unsharded fused kernel on the whole table

Prints one JSON line (rank 0): samples/s over all ranks (max-over-ranks device time), the
achieved HBM GB/s per GPU, and both transports when --transport both.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def jagged_keys(rank, B, max_len, batch_index=0):
    from recommendflow_b200.synth import decimal_keys
    rng = np.random.default_rng(4242 + rank + 1000 * batch_index)
    lens = rng.integers(1, max_len + 1, size=B)
    bag = np.zeros(B + 1, dtype=np.int32)
    bag[1:] = np.cumsum(lens)
    v = rng.integers(0, 10**9, size=int(bag[-1]))
    arena, offs = decimal_keys(b"item_", v)
    return arena, offs, bag


def parse(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100_000_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--max-len", type=int, default=200)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--transport", default="both", choices=["p2p", "nccl", "both"])
    ap.add_argument("--graph", action="store_true", help="capture each step (p2p transport) in a CUDA graph and replay")
    return ap.parse_args(argv)


def run(args, world, rank, dev):
    """Measure C4 on an already initialised process group; returns the result dict (rank 0) or None."""
    import torch
    import torch.distributed as dist
    from recommendflow_b200 import _native as nat
    from recommendflow_b200.bag_ops import FieldCall, bag_forward
    from recommendflow_b200.sharded import ShardedEmbeddingBag
    from recommendflow_b200.strings import StringColumn

    B, N, D, K, W = args.batch, args.rows, args.dim, args.steps, max(args.warmup, 3)
    NB = 4
    batches = []
    for bi in range(NB):
        arena, offs, bag = jagged_keys(rank, B, args.max_len, bi)
        batches.append(StringColumn.from_arena(arena, offs, (B, None), bag).to(dev))
    max_keys = max(c.n_items for c in batches)
    mean_len = float(np.mean([c.n_items for c in batches])) / B
    key_bytes = float(np.mean([c.nbytes for c in batches])) / B
    # algorithmic bytes per sample (bag): rows + key bytes + offsets + output
    bps = mean_len * 4 * D + key_bytes + mean_len * 4 + 4 * D

    def timed(step_fn):
        for i in range(W):
            step_fn(i)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            step_fn(W + i)
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    results = {}
    if world == 1:
        table = torch.empty(N, D, dtype=torch.float32, device=dev).uniform_(-0.05, 0.05)
        out = torch.empty(B, D, dtype=torch.float32, device=dev)
        calls = [[FieldCall([(table, N, None)], D, "avg", keys=c, mask_mode=nat.MASK_EMPTY_STRING, out=out)] for c in batches]
        ms = timed(lambda i: bag_forward(calls[i % NB], B))
        results["single_gpu_fused"] = ms
    else:
        transports = ["p2p", "nccl"] if args.transport == "both" else [args.transport]
        outs, shard = {}, None
        for tr in transports:
            layer = ShardedEmbeddingBag(N, D, combiner="avg", salt=None, mask_value="", transport=tr, max_batch=B,
                                        max_keys=max_keys)
            if shard is None:
                shard = layer.shard
            else:
                layer.shard = shard                     # same table for both transports
            out = torch.empty(B, D, dtype=torch.float32, device=dev)
            ms = timed(lambda i: layer(batches[i % NB], out=out))
            results[tr] = ms
            if tr == "p2p":
                # steady-state pipelining: routing of step i+1 (side stream) overlaps pooling of step i
                state = {"ticket": None}

                def piped(i):
                    if state["ticket"] is None:
                        state["ticket"] = layer.prepare(batches[i % NB])
                    nxt = layer.prepare(batches[(i + 1) % NB])
                    layer.finish(state["ticket"], out=out)
                    state["ticket"] = nxt
                results["p2p_pipelined"] = timed(piped)
                layer.finish(state["ticket"], out=out)          # drain
                torch.cuda.synchronize()
                results["p2p_pipelined_equals_eager"] = bool(torch.equal(out, layer(batches[(W + K) % NB]).clone()))
            if tr == "p2p" and args.graph:
                # the p2p step has no host synchronisation, so the whole step (route kernels, symmetric-
                # memory barriers, fused gather+pool into peer memory, combine) replays as one graph launch
                graphs = []
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for c in batches:
                        layer(c, out=out)
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                dist.barrier()
                for c in batches:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        layer(c, out=out)
                    graphs.append(g)
                results["p2p_graph"] = timed(lambda i: graphs[i % NB].replay())
                results["p2p_graph_equals_eager"] = bool(torch.equal(out, layer(batches[(W + K - 1) % NB]).clone()))
            outs[tr] = layer(batches[0]).clone()
            if tr == "p2p":                             # per-phase device times of a few steps (rank 0 reports)
                acc = {}
                for i in range(5):
                    layer.profile = []
                    layer(batches[i % NB], out=out)
                    torch.cuda.synchronize()
                    for (n0, e0), (n1, e1) in zip(layer.profile[:-1], layer.profile[1:]):
                        acc[n1] = acc.get(n1, 0.0) + e0.elapsed_time(e1) / 5
                layer.profile = None
                results["p2p_phases_ms"] = {k: round(v, 4) for k, v in acc.items()}
            del layer
        if len(outs) == 2:
            results["p2p_equals_nccl"] = bool(torch.equal(outs["p2p"], outs["nccl"]))
    line = None
    if rank == 0:
        best = min(v for k, v in results.items() if isinstance(v, float))
        line = {"metric": "lookup+pool samples/sec", "workload": f"c4: jagged 1..{args.max_len} keys/bag (mean {mean_len:.1f}), avg pooling, "
                          f"{N}-row x {D}-dim fp32 table row-sharded id % {world}, batch {B}/GPU",
                "n_gpus": world, "value": world * B / (best / 1e3), "unit": "samples/s", "ms_per_step": results,
                "steps": K, "warmup": W, "algorithmic_bytes_per_sample": bps,
                "hbm_gbs_per_gpu": bps * B / (best / 1e3) / 1e9, "gpu_launches": nat.launch_count()}
    return line


def main():
    import torch
    import torch.distributed as dist
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    line = run(args, world, rank, dev)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
