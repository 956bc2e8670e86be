#!/usr/bin/env python
"""ONE-GPU emulation of one rank's work in the W-way row-sharded C4 step.

Everything a rank does per step is local compute except the NVLink hops and the two barriers:
  source role : route its B bags to W owners                     (rf_shard_route_tiles, W destination buffers)
  owner role  : pool the rows W sources asked of its shard        (rf_bag_forward over W gapped (rows, begin, end) lists)
  source role : combine W partials / finish the accumulator      (rf_combine_partials)
so with all "peer" buffers local the kernels see the real per-rank key counts, table shard size and
access pattern.  This is where the route kernel, the co-residency of route(i+1) with pool(i) and the
combine-free (red.global.add) mode are tuned for one GPU-minute instead of eight; the N-GPU numbers come
from tools/bench_sharded.py.

    python tools/emu_sharded.py [--world 8] [--batch 65536] [--prehashed] [--steps 20]
Prints one JSON line.  RF_ROUTE_STAGE=0 selects the round-1 route kernel (A/B in separate processes).
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--rows", type=int, default=100_000_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--max-len", type=int, default=200)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--prehashed", action="store_true")
    ap.add_argument("--pool-ctas", type=int, nargs="*", default=[0, 3])
    ap.add_argument("--label", default="")
    args = ap.parse_args()

    import torch
    from recommendflow_b200 import _native as nat
    from recommendflow_b200.bag_ops import hash_strings
    from recommendflow_b200.sharded import CudaShardOps, shard_rows
    from recommendflow_b200.strings import StringColumn
    from tools.bench_sharded import jagged_keys

    W, B, N, D, K = args.world, args.batch, args.rows, args.dim, args.steps
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    ops = CudaShardOps()
    shard = torch.empty(shard_rows(N, 0, W), D, dtype=torch.float32, device=dev).uniform_(-0.05, 0.05)

    # ---- the W sources' batches; owner 0 keeps what each source routes to it -------------------------------
    cols = []
    for s in range(W):
        arena, offs, bag = jagged_keys(s, B, args.max_len)
        cols.append(StringColumn.from_arena(arena, offs, (B, None), bag).to(dev))
    max_keys = max(c.n_items for c in cols)
    ids = [hash_strings(c, N, "", None).view(-1) for c in cols] if args.prehashed else None
    rows0 = torch.zeros(W, max_keys, dtype=torch.int64, device=dev)          # [source] -> rows for owner 0
    beg0 = torch.zeros(W, B, dtype=torch.int32, device=dev)
    end0 = torch.zeros(W, B, dtype=torch.int32, device=dev)
    rows_all = torch.zeros(W, max_keys, dtype=torch.int64, device=dev)       # the timed source's W destinations
    beg_all = torch.zeros(W, B, dtype=torch.int32, device=dev)
    end_all = torch.zeros(W, B, dtype=torch.int32, device=dev)
    ids_ws = torch.empty(max_keys, dtype=torch.int64, device=dev)

    def route(s, rows_dst, beg_dst, end_dst, ws=ids_ws):
        keys = ids[s] if args.prehashed else cols[s]
        ops.route_tiles(keys, N, "", None, ws, cols[s].bag_offsets, 0, B, W, rows_dst, beg_dst, end_dst)

    for s in range(W):          # owner 0's view: destination 0 is the real buffer, the others share scratch
        route(s, [rows0[s].data_ptr()] + [rows_all[g].data_ptr() for g in range(1, W)],
              [beg0[s].data_ptr()] + [beg_all[g].data_ptr() for g in range(1, W)],
              [end0[s].data_ptr()] + [end_all[g].data_ptr() for g in range(1, W)])
    torch.cuda.synchronize()
    keys_owner0 = int((end0 - beg0).sum().item())

    partials = torch.empty(W, B, D, dtype=torch.float32, device=dev)
    acc = torch.zeros(1, B, D, dtype=torch.float32, device=dev)
    out = torch.empty(B, D, dtype=torch.float32, device=dev)
    est = max(1, max_keys // W)

    def route_step(i):
        s = i % W
        route(s, [rows_all[g].data_ptr() for g in range(W)], [beg_all[g].data_ptr() for g in range(W)],
              [end_all[g].data_ptr() for g in range(W)])

    def pool_step(accumulate, ctas):
        if accumulate:
            acc.zero_()
        ops.pool(shard, [rows0[s] for s in range(W)], [beg0[s] for s in range(W)],
                 [acc[0] if accumulate else partials[s] for s in range(W)], B, "sum", est, [end0[s] for s in range(W)],
                 accumulate=accumulate, max_ctas_per_sm=ctas)

    def combine_step(accumulate):
        ops.combine(acc if accumulate else partials, 1 if accumulate else W, B, D, "avg", 0, cols[0].bag_offsets, out)

    def timed(fn, n=K, warm=3):
        for i in range(warm):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(warm + i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    res = {"label": args.label, "world": W, "batch": B, "keys_per_source": max_keys, "keys_owner0": keys_owner0,
           "prehashed": args.prehashed, "route_stage_env": os.environ.get("RF_ROUTE_STAGE", "1")}
    res["route_ms"] = timed(route_step)
    # parity of the two pooling modes (same inputs): ordered partials + combine vs accumulate + finish
    pool_step(False, 0)
    combine_step(False)
    ordered = out.clone()
    pool_step(True, 0)
    combine_step(True)
    torch.cuda.synchronize()
    res["acc_vs_ordered_max_abs"] = float((out - ordered).abs().max().item())
    for accumulate in (False, True):
        tag = "acc" if accumulate else "ordered"
        res[f"pool_{tag}_ms"] = timed(lambda i: pool_step(accumulate, 0))
        res[f"combine_{tag}_ms"] = timed(lambda i: combine_step(accumulate))
        res[f"serial_{tag}_ms"] = timed(lambda i: (route_step(i), pool_step(accumulate, 0), combine_step(accumulate)))
        # pipelined: route of the next step on a high-priority side stream under this step's pool
        for ctas in args.pool_ctas:
            sR = torch.cuda.Stream(device=dev, priority=-1)
            sC = torch.cuda.Stream(device=dev, priority=-1)
            cur = torch.cuda.current_stream(dev)

            def piped(i):
                ev0 = torch.cuda.Event()
                ev0.record(cur)
                with torch.cuda.stream(sR):
                    sR.wait_event(ev0)
                    route_step(i + 1)
                    r_done = torch.cuda.Event()
                    r_done.record(sR)
                pool_step(accumulate, ctas)
                p_done = torch.cuda.Event()
                p_done.record(cur)
                with torch.cuda.stream(sC):
                    sC.wait_event(p_done)
                    combine_step(accumulate)
                cur.wait_event(r_done)      # the next pool needs this routing; the combine overlaps it (as in the
                                            # real pipeline, where the exchange buffers are double-buffered)

            res[f"piped_{tag}_ctas{ctas}_ms"] = timed(piped)
    res["gpu_launches"] = nat.launch_count()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
