#!/usr/bin/env python
"""Single-GPU microbenchmark of the routing kernels (rf_shard_route_keys) with `world` virtual
owners whose buffers are all local -- lets ncu see count / scan / scatter separately."""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--max-len", type=int, default=200)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    import torch
    from recommendflow_b200 import _native as nat
    from recommendflow_b200.sharded import CudaShardOps
    from recommendflow_b200.strings import StringColumn
    from tools.bench_sharded import jagged_keys

    B, W = args.batch, args.world
    arena, offs, bag = jagged_keys(0, B, args.max_len)
    col = StringColumn.from_arena(arena, offs, (B, None), bag).to("cuda")
    n = col.n_items
    ops = CudaShardOps()
    ids_ws = torch.empty(n, dtype=torch.int64, device="cuda")
    counts = torch.empty(W * B, dtype=torch.int32, device="cuda")
    offs_local = torch.empty(W * (B + 1 + (B + 1023) // 1024), dtype=torch.int32, device="cuda")
    rows = torch.empty(W, n, dtype=torch.int64, device="cuda")
    offs_dst = torch.empty(W, B + 1, dtype=torch.int32, device="cuda")

    def step():
        ops.route_keys(col, 100_000_000, "", None, ids_ws, col.bag_offsets, 0, B, W, counts, offs_local,
                       [offs_dst[g].data_ptr() for g in range(W)], [rows[g].data_ptr() for g in range(W)])

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    print(f"route_keys: B={B} world={W} keys={n}: {e0.elapsed_time(e1) / args.steps * 1e3:.1f} us/step")
    # sanity: CSR totals add up to the key count
    assert int(offs_dst[:, B].sum()) == n
    # single-pass tile routing
    begs = torch.empty(W, B, dtype=torch.int32, device="cuda")
    ends = torch.empty(W, B, dtype=torch.int32, device="cuda")

    def step2():
        ops.route_tiles(col, 100_000_000, "", None, ids_ws, col.bag_offsets, 0, B, W, [rows[g].data_ptr() for g in range(W)],
                        [begs[g].data_ptr() for g in range(W)], [ends[g].data_ptr() for g in range(W)])

    for _ in range(3):
        step2()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        step2()
    e1.record()
    torch.cuda.synchronize()
    print(f"route_tiles: B={B} world={W} keys={n}: {e0.elapsed_time(e1) / args.steps * 1e3:.1f} us/step")
    assert int((ends - begs).sum()) == n


if __name__ == "__main__":
    main()
