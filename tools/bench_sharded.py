#!/usr/bin/env python
"""C4: jagged behaviour sequences (1..max_len keys), mean pooling, one
rows x dim fp32 table row-sharded (id % world) across the GPUs of one box.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_sharded.py [...]
    python tools/bench_sharded.py            # N = 1: the unsharded fused kernel on the whole table

Prints one JSON line (rank 0).  Per batch size (65 536 and SURVEY's 8 192 per GPU) and key kind (string keys hashed
on the fly, pre-hashed int64 ids): step time of every variant of the p2p step (ordered partials + combine, or the
combine-free accumulate mode; eager / pipelined / CUDA graph), samples/s over all ranks (max-over-ranks device
time), the same-box single-GPU time measured in the same run (rank 0 builds the whole table once) and
`speedup_vs_n1`, and `parity_check`: every rank compares its first bags with the CPU oracle on rows rebuilt from
the table's closed form -- bit-exact for the ordered path, the fp32 re-association bound for the accumulate path.
A parity failure ends the run with a non-zero exit code.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

N_CHECK_BAGS = 512


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
    ap.add_argument("--batch", type=int, nargs="*", default=[65536, 8192])
    ap.add_argument("--max-len", type=int, default=200)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--nccl", action="store_true", help="also time the all_to_all (NCCL) transport baseline")
    ap.add_argument("--no-n1", action="store_true", help="skip the same-box single-GPU measurement on rank 0")
    ap.add_argument("--quick", action="store_true", help="string keys at the first batch size only")
    ap.add_argument("--no-train", action="store_true", help="skip the C5 training-step measurement (N > 1)")
    return ap.parse_args(argv)


# ------------------------------------------------------------------------------------------------------------
# the checker: CPU oracle on rows rebuilt from the closed form (tools / tests side only)
# ------------------------------------------------------------------------------------------------------------
def host_reference(ids, bag, world, dim, n_bags, ordered):
    """ids: bucket ids of this rank's keys (oracle), bag: CSR.  Mean pooling of the first n_bags bags.
    ordered=True : the sharded algorithm restated -- fp32 partial sums per owner in key order, combined in rank order,
                   one fp32 divide (bit-exact target);  world == 1 is the plain sequential sum of the fused kernel.
    ordered=False: float64 sums (target of the re-association bound)."""
    from recommendflow_b200.synth import closed_form_rows
    ids = np.asarray(ids[:bag[n_bags]], dtype=np.int64)
    uniq, inv = np.unique(ids, return_inverse=True)
    rows = closed_form_rows(uniq, dim)
    out = np.zeros((n_bags, dim), dtype=np.float32 if ordered else np.float64)
    for b in range(n_bags):
        lo, hi = int(bag[b]), int(bag[b + 1])
        if hi == lo:
            continue
        if not ordered:
            out[b] = rows[inv[lo:hi]].astype(np.float64).sum(axis=0) / (hi - lo)
            continue
        owners = ids[lo:hi] % world
        total = None
        for g in range(world):
            acc = np.zeros(dim, dtype=np.float32)
            for k in np.nonzero(owners == g)[0]:
                acc = acc + rows[inv[lo + k]]
            total = acc if total is None else total + acc
        out[b] = total / np.float32(hi - lo)
    return out


def check(got, ids, bag, world, dim, ordered, max_len):
    n = min(N_CHECK_BAGS, got.shape[0])
    want = host_reference(ids, bag, world, dim, n, ordered)
    g = got[:n].cpu().numpy()
    if ordered:
        ok = bool(np.array_equal(g.view(np.uint32), want.view(np.uint32)))
        return ok, 0.0 if ok else float(np.abs(g - want).max())
    err = float(np.abs(g.astype(np.float64) - want).max())
    return err <= max_len * 0.05 * 2.0 ** -21, err


def run_train_c5(world, rank, dev, barrier, steps=5):
    """C5 (BASELINE.json configs[4]): full training step, 1 B table rows in total row-sharded over the ranks, dense
    towers data-parallel, in-batch softmax over the global batch (recommendflow_b200/training_sharded.py).
    4 features x 250 M rows x 16 dims (64 GB of tables + 128 GB of Adam moments over the box), Keras Adam on every row
    (the reference's optimizer semantics), batch 8192 per GPU (the CUDA-core loss backward bounds the step today;
    65 536 per GPU needs the tensor-core backward)."""
    import torch
    import torch.distributed as dist
    from recommendflow_b200.backend.blocks.mlp import create_mlp
    from recommendflow_b200.sharded import ShardedEmbeddingBag
    from recommendflow_b200.strings import StringColumn
    from recommendflow_b200.training_sharded import ShardedRecallTrainer
    rows_total, n_feat, D, B, max_len = 1_000_000_000, 4, 16, 8192, 20
    N = rows_total // n_feat
    if N // world * D * 4 * 3 * n_feat > 150e9:
        return {"skipped": f"1 B rows need more than {world} GPUs for tables + Adam moments"}
    names = [f"user_{i}" for i in range(n_feat // 2)] + [f"ad_{i}" for i in range(n_feat // 2)]
    batches = []
    for bi in range(2):
        b = {}
        for i, n in enumerate(names):
            arena, offs, bag = jagged_keys(rank + 100 * i, B, max_len, bi)
            b[n] = StringColumn.from_arena(arena, offs, (B, None), bag).to(dev)
        batches.append(b)
    y = torch.ones(B, device=dev)

    def one(transport):
        bags = {n: ShardedEmbeddingBag(N, D, combiner="avg", salt=None, mask_value="", transport=transport, max_batch=B,
                                       max_keys=B * max_len) for n in names}
        torch.manual_seed(11)
        towers = [create_mlp([256, 128], 0.0, "selu", None, name=t) for t in ("user_tower", "ad_tower")]
        x = torch.zeros(2, D * n_feat // 2)
        for t in towers:
            t(x)
            t.to(dev)
        trainer = ShardedRecallTrainer({n: bags[n] for n in names[:n_feat // 2]}, {n: bags[n] for n in names[n_feat // 2:]},
                                       towers[0], towers[1], learning_rate=1e-3)
        losses = [float(trainer.train_step(batches[i % 2], y)) for i in range(2)]          # warm-up (builds optimizers)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            losses.append(float(trainer.train_step(batches[i % 2], y)))
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        del trainer, bags
        torch.cuda.empty_cache()
        return float(t.item()), losses

    # both exchanges of the sharded bags: nccl (all_to_all forward, all_gather backward) and p2p (routing / pooled vectors /
    # gradients through NVLink peer memory, device-side barriers)
    per_transport = {}
    for transport in ("nccl", "p2p"):
        try:
            per_transport[transport] = one(transport)
        except Exception as exc:
            per_transport[transport] = (None, f"{type(exc).__name__}: {exc}")
    timed = {k: v for k, v in per_transport.items() if v[0] is not None}
    if not timed:
        return {"error": {k: v[1] for k, v in per_transport.items()}}
    best = min(timed, key=lambda k: timed[k][0])
    ms, losses = timed[best]
    return {"workload": f"c5: training step, {n_feat} features x {N} rows x {D} dims = {rows_total} table rows row-sharded id % {world} "
                        f"(Keras Adam on every row), towers [256, 128] data-parallel, in-batch softmax over the global batch "
                        f"{world} x {B}, jagged 1..{max_len} keys/bag",
            "ms_per_step": ms, "samples_per_s": world * B / (ms / 1e3), "steps": steps, "transport": best,
            "ms_per_step_by_transport": {k: (v[0] if v[0] is not None else v[1]) for k, v in per_transport.items()},
            "losses": [round(v, 5) for v in losses], "loss_fell": losses[-1] < losses[0]}


def run(args, world, rank, dev):
    """Measure C4 on an already initialised process group; returns the result dict (rank 0) or None."""
    import torch
    import torch.distributed as dist
    import oracle
    from recommendflow_b200 import _native as nat
    from recommendflow_b200.bag_ops import BagPlan, FieldCall, hash_strings
    from recommendflow_b200.sharded import BucketIds, ShardedEmbeddingBag
    from recommendflow_b200.strings import StringColumn
    from recommendflow_b200.synth import fill_closed_form

    N, D, K, W = args.rows, args.dim, args.steps, max(args.warmup, 3)
    NB = 4
    batch_sizes = args.batch[:1] if args.quick else args.batch
    kinds = ["string"] if args.quick else ["string", "prehashed"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn):
        for i in range(W):
            step_fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            step_fn(W + i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / K
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- inputs: per batch size, NB rotating batches; ids from the GPU hash (checked against the oracle below) ----
    data = {}
    for B in batch_sizes:
        cols, host = [], []
        for bi in range(NB):
            arena, offs, bag = jagged_keys(rank, B, args.max_len, bi)
            cols.append(StringColumn.from_arena(arena, offs, (B, None), bag).to(dev))
            if bi == 0:
                nk = int(bag[min(N_CHECK_BAGS, B)])
                host = (oracle.hash_strings(arena[:offs[nk]], offs[:nk + 1], N, "", None), bag)
        ids = [BucketIds(hash_strings(c, N, "", None).view(-1), c.bag_offsets) for c in cols]
        assert np.array_equal(ids[0].ids[:len(host[0])].cpu().numpy(), host[0]), "GPU bucket ids differ from the oracle's"
        mean_len = float(np.mean([c.n_items for c in cols])) / B
        key_bytes = float(np.mean([c.nbytes for c in cols])) / B
        data[B] = {"string": cols, "prehashed": ids, "host": host, "max_keys": max(c.n_items for c in cols),
                   # algorithmic bytes per sample: rows + keys (+ offsets) + output
                   "bps": {"string": mean_len * 4 * D + key_bytes + mean_len * 4 + 4 * D,
                           "prehashed": mean_len * 4 * D + mean_len * 8 + 4 * D}, "mean_len": mean_len}

    results, parity, launches0 = {}, {}, nat.launch_count()

    def single_gpu(tag_prefix):
        """The unsharded fused kernel on the whole table (this rank only)."""
        table = fill_closed_form(torch.empty(N, D, dtype=torch.float32, device=dev))
        for B in batch_sizes:
            out = torch.empty(B, D, dtype=torch.float32, device=dev)
            for kind in kinds:
                if kind == "string":
                    plans = [BagPlan([FieldCall([(table, N, None)], D, "avg", keys=c, mask_mode=nat.MASK_EMPTY_STRING, out=out)], B)
                             for c in data[B]["string"]]
                else:
                    plans = [BagPlan([FieldCall([(table, N, None)], D, "avg", ids=i.ids.view(1, -1), bag_offsets=i.bag_offsets,
                                                out=out, n_items=i.n_items)], B) for i in data[B]["prehashed"]]
                for i in range(W):
                    plans[i % NB].launch()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(K):
                    plans[(W + i) % NB].launch()
                e1.record()
                torch.cuda.synchronize()
                results[f"{tag_prefix}/B{B}/{kind}"] = e0.elapsed_time(e1) / K
                plans[0].launch()
                torch.cuda.synchronize()
                ok, err = check(out, data[B]["host"][0], data[B]["host"][1], 1, D, True, args.max_len)
                parity[f"{tag_prefix}/B{B}/{kind}"] = {"bit_exact_vs_oracle": ok, "max_abs_err": err}
        del table
        torch.cuda.empty_cache()

    if world == 1:
        single_gpu("n1")
    else:
        if not args.no_n1:          # the same box's single-GPU time, for speedup_vs_n1 (rank 0 holds the 51 GB table once)
            if rank == 0:
                single_gpu("n1")
            barrier()
        shard = None
        for B in batch_sizes:
            d = data[B]
            for mode in ("ordered", "accumulate"):
                layer = ShardedEmbeddingBag(N, D, combiner="avg", salt=None, mask_value="", transport="p2p", max_batch=B,
                                            max_keys=d["max_keys"], deterministic=(mode == "ordered"), keep_ids=False)
                if shard is None:
                    shard = fill_closed_form(layer.shard.data, first_row=rank, row_stride=world)
                else:
                    layer.shard = torch.nn.Parameter(shard, requires_grad=False)
                out = torch.empty(B, D, dtype=torch.float32, device=dev)
                for kind in kinds:
                    batches = d[kind]
                    tag = f"p2p/B{B}/{kind}/{mode}"
                    if B >= 32768:
                        results[tag + "/eager"] = timed(lambda i: layer(batches[i % NB], out=out))
                        state = {"ticket": None}

                        def piped(i):
                            if state["ticket"] is None:
                                state["ticket"] = layer.prepare(batches[i % NB])
                            nxt = layer.prepare(batches[(i + 1) % NB])
                            layer.finish(state["ticket"], out=out)
                            state["ticket"] = nxt
                        results[tag + "/pipelined"] = timed(piped)
                        layer.finish(state["ticket"], out=out)          # drain
                        torch.cuda.synchronize()
                    else:
                        # small batches are launch-bound: the whole PIPELINED step -- route + barrier of batch i+1 on one
                        # branch, pool / drain / combine of batch i on the other; no host sync anywhere -- is recorded
                        # once per rotating batch and replayed as one graph launch
                        results[tag + "/eager"] = timed(lambda i: layer(batches[i % NB], out=out))
                        barrier()
                        tix = [layer.prepare(batches[0])]                # tix[i]: the ticket graph i finishes
                        barrier()
                        graphs = []
                        for i in range(NB):
                            g = torch.cuda.CUDAGraph()
                            with torch.cuda.graph(g):
                                nxt = layer.prepare(batches[(i + 1) % NB], capturing=True)
                                layer.finish(tix[i], out=out)
                            graphs.append(g)
                            tix.append(nxt)
                        results[tag + "/pipelined_graph"] = timed(lambda i: graphs[i % NB].replay())
                        # replay W + K - 1 was the last: `out` holds batch (W + K - 1) % NB, batch (W + K) % NB is routed
                        piped_out = out.clone()
                        layer.finish(dict(tix[(W + K) % NB], overlap=False, capturing=False), out=out)     # drain, eagerly
                        torch.cuda.synchronize()
                        del graphs
                        nat.lib().rf_release_captured_launches()
                        eager_out = layer(batches[(W + K - 1) % NB], out=out)
                        torch.cuda.synchronize()
                        if mode == "ordered" and not torch.equal(piped_out, eager_out):
                            raise SystemExit(f"bench_sharded: {tag}: graph-replayed pipelined step differs from the eager step")
                    # ---- parity of THIS variant against the oracle (every rank checks its own first bags) ----
                    got = layer(batches[0], out=out)
                    torch.cuda.synchronize()
                    ok, err = check(got, d["host"][0], d["host"][1], world, D, mode == "ordered", args.max_len)
                    flags = torch.tensor([0 if ok else 1], dtype=torch.int32, device=dev)
                    errs = torch.tensor([err], dtype=torch.float64, device=dev)
                    dist.all_reduce(flags, op=dist.ReduceOp.SUM)
                    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
                    parity[tag] = {("bit_exact_vs_oracle" if mode == "ordered" else "within_reassociation_bound"):
                                   int(flags.item()) == 0, "ranks_checked": world, "max_abs_err": float(errs.item())}
                    if mode == "ordered" and kind == "string":
                        acc = {}
                        for i in range(5):
                            layer.profile = []
                            layer(batches[i % NB], out=out)
                            torch.cuda.synchronize()
                            for (n0, e0), (n1, e1) in zip(layer.profile[:-1], layer.profile[1:]):
                                acc[n1] = acc.get(n1, 0.0) + e0.elapsed_time(e1) / 5
                        layer.profile = None
                        results["phases_ms/" + tag] = {k: round(v, 4) for k, v in acc.items()}
                del layer
            if args.nccl:
                layer = ShardedEmbeddingBag(N, D, combiner="avg", salt=None, mask_value="", transport="nccl", max_batch=B,
                                            max_keys=d["max_keys"])
                layer.shard = torch.nn.Parameter(shard, requires_grad=False)
                out = torch.empty(B, D, dtype=torch.float32, device=dev)
                results[f"nccl/B{B}/string"] = timed(lambda i: layer(d["string"][i % NB], out=out))
                del layer

    train_c5 = None
    if world > 1 and not args.no_train and not args.quick:
        try:
            train_c5 = run_train_c5(world, rank, dev, barrier)
        except Exception as exc:                      # the forward measurement above must survive a training-side failure
            train_c5 = {"error": f"{type(exc).__name__}: {exc}"}

    line = None
    bad = [k for k, v in parity.items() if not all(x for x in v.values() if isinstance(x, bool))]
    if rank == 0:
        summary = {}
        for B in batch_sizes:
            for kind in kinds:
                n1 = results.get(f"n1/B{B}/{kind}")
                entry = {"n1_ms": n1, "algorithmic_bytes_per_sample": data[B]["bps"][kind]}
                if world == 1:
                    entry.update(best_ms=n1, best_variant="single_gpu_fused", samples_per_s=B / (n1 / 1e3),
                                 hbm_gbs_per_gpu=data[B]["bps"][kind] * B / (n1 / 1e3) / 1e9)
                else:
                    cand = {k: v for k, v in results.items() if k.startswith(f"p2p/B{B}/{kind}/") and isinstance(v, float)}
                    det = {k: v for k, v in cand.items() if "/ordered/" in k}
                    best_k = min(cand, key=cand.get)
                    best_d = min(det, key=det.get)
                    entry.update(best_ms=cand[best_k], best_variant=best_k, samples_per_s=world * B / (cand[best_k] / 1e3),
                                 hbm_gbs_per_gpu=data[B]["bps"][kind] * B / (cand[best_k] / 1e3) / 1e9,
                                 best_ordered_ms=det[best_d], best_ordered_variant=best_d)
                    if n1:
                        entry["speedup_vs_n1"] = world * n1 / cand[best_k]
                        entry["speedup_vs_n1_ordered"] = world * n1 / det[best_d]
                summary[f"B{B}/{kind}"] = entry
        head = summary[f"B{batch_sizes[0]}/string"]
        limiter = None
        ph = results.get(f"phases_ms/p2p/B{batch_sizes[0]}/string/ordered")
        if ph:
            top = max(ph, key=ph.get)
            limiter = (f"eager step on rank 0: {ph}; largest phase '{top}' (the fused gather+pool, HBM-bound); what is left "
                       f"beside it: route + barriers + combine = {sum(v for k, v in ph.items() if k != 'pool'):.3f} ms")
        line = {"metric": "lookup+pool samples/sec",
                "workload": f"c4: jagged 1..{args.max_len} keys/bag (mean {data[batch_sizes[0]]['mean_len']:.1f}), avg pooling, "
                            f"{N}-row x {D}-dim fp32 table row-sharded id % {world}, batch per GPU {batch_sizes}, "
                            f"string keys (Fingerprint64 on the fly) and pre-hashed int64 ids",
                "n_gpus": world, "value": head["samples_per_s"], "unit": "samples/s", "summary": summary,
                "speedup_vs_n1": head.get("speedup_vs_n1"), "ms_per_step": results, "steps": K, "warmup": W,
                "parity_check": parity, "parity_ok": not bad, "limiter": limiter, "train_c5": train_c5,
                "gpu_launches": nat.launch_count() - launches0}
    if bad:
        if rank == 0:
            print(json.dumps(line), file=sys.stderr)
        raise SystemExit(f"bench_sharded: parity FAILED for {bad}")
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
