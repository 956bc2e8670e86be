#!/usr/bin/env python
"""bench.py -- lookup+pool samples/s of the fused hash + gather + pool path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c2x2|c2zipf|c2l1|c3|c1|small]

Workload (config.workload) at every N: SURVEY.md §8(d) "C2" = BASELINE.json configs[1]: 26 hashed
sparse fields, 1M-row x 64-dim fp32 tables, batch 65536, sum pooling, 4 keys per bag, keys
"fNN_<v>" (FarmHash Fingerprint64 mod N, Keras mask_value=""), tables U(-0.05, 0.05).
A step = one pass of the hot path over one batch = ONE fused kernel launch.
N > 1: one process per GPU (torchrun), each rank holds a full replica of the 6.66 GB tables and
its own batch (what the reference's MirroredStrategy does) -- no data-path collective, weak scaling.

`--impl reference` times the CPU restatement of the reference's TF path (oracle/, all host
threads): TensorFlow is not installable here, so the oracle port is the reference arm.
"""
import argparse
import json
import os
import sys
import threading
import time

# the CPU arms run OpenMP with a fixed thread count: idle workers must sleep, not spin, or a box that grants
# fewer cores than it shows makes the baseline collapse (measured here: 4.7 ms on 1 thread vs 98 ms on 4 spinning)
os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "lookup+pool samples/sec"
UNIT = "samples/s"

WORKLOADS = {
    #        fields rows     dim  batch  L  tables/field
    "c2":   (26, 1_000_000, 64, 65536, 4, 1),
    "c2x2": (26, 1_000_000, 64, 65536, 4, 2),     # reference-faithful double SipHash variant
    "c2zipf": (26, 1_000_000, 64, 65536, 4, 1),   # SURVEY.md §8(d) secondary run: key values ~ Zipf(1.05) (hot rows)
    "c2l1": (26, 1_000_000, 64, 65536, 1, 1),     # SURVEY.md §8(d) corner run: one key per bag
    # C3's embedding side: the normalised base_recall_sdpa plan -- 228 hashed fields x 2 tables of 100000 x 8
    "c3":   (228, 100_000, 8, 8192, 1, 2),
    "small": (4, 3000, 16, 2048, 4, 2),
    # SURVEY.md §8(d) C1 substitute: the one hashing feature of conf/base_conf.yaml (app_id: N = 3000, D = 16, sum,
    # seeds [2022, 2023]); the reference's demo_conf.yaml model itself cannot run (needs TF + bert4keras + a checkpoint)
    "c1": (1, 3000, 16, 8192, 4, 2),
}
KEY_ZIPF = {"c2zipf": 1.05}
N_KEY_BATCHES = 4
E2E_CHUNKS = 8
E2E_STREAMS = int(os.environ.get("RF_E2E_STREAMS", "4"))    # a chunk's H2D + kernel must never leave the D2H engine idle


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-c4", action="store_true", help="skip the row-sharded C4 measurement appended to the line")
    ap.add_argument("--no-c3", action="store_true", help="skip the C3 end-to-end forward appended to the line (N = 1)")
    ap.add_argument("--no-train", action="store_true", help="skip the C3 training step appended to the line (N = 1)")
    return ap.parse_args()


def workload_config(name, n_gpus=1):
    F, N, D, B, L, T = WORKLOADS[name]
    return {"parallelism": f"dp{n_gpus} (replicated tables, one batch per GPU, no data-path collective)",
            "workload": f"{name}: {F} hashed fields x {T} table(s), {N}-row x {D}-dim fp32 tables, batch {B}, "
                        f"{L} keys/bag, sum pooling, {'SipHash-2-4 seeds [2022,2023]' if T == 2 else 'Fingerprint64'}"
                        f" mod N, mask_value='', key values ~ "
                        f"{'Zipf(%g)' % KEY_ZIPF[name] if name in KEY_ZIPF else 'Uniform[0, 1e7)'}",
            "fields": F, "rows": N, "dim": D, "batch_per_gpu": B, "bag_len": L, "tables_per_field": T,
            "l2_policy": (f"inputs larger than L2: {F * T * N * D * 4 / 1e9:.2f} GB of tables gathered at random rows; "
                          f"{N_KEY_BATCHES} distinct key batches rotate across steps") if F * T * N * D * 4 > 4 * 126e6 else
                         (f"tables ({F * T * N * D * 4 / 1e6:.1f} MB) fit in the 126 MB L2 and stay resident across steps: this "
                          f"workload measures launch, hashing and output traffic, not table reads from HBM")}


def salts_for(T):
    return [None] if T == 1 else [[2022, 2022], [2023, 2023]]


def make_keys(name, rank):
    """{field name: (arena, offsets, shape)} for each of the rotating key batches."""
    from recommendflow_b200.synth import c2_field_keys
    F, N, D, B, L, T = WORKLOADS[name]
    batches = []
    for bi in range(N_KEY_BATCHES):
        fields = {}
        for f in range(F):
            arena, offs = c2_field_keys(f, B, L, batch_index=bi + 16 * rank, zipf=KEY_ZIPF.get(name))
            fields[f"f{f:02d}"] = (arena, offs, (B, L))
        batches.append(fields)
    return batches


def algorithmic_bytes_per_sample(name, key_batches):
    """sum_f [T*l*4D (row gathers) + l*(s + 4) (key bytes + offsets) + T*4D (output)]."""
    F, N, D, B, L, T = WORKLOADS[name]
    key_bytes = np.mean([sum(a.size for a, _, _ in kb.values()) for kb in key_batches]) / B
    return F * (T * L * 4 * D + L * 4 + T * 4 * D) + key_bytes


# ------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------
class ClockSampler(object):
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.001)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return None
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's TF CPU path
# ------------------------------------------------------------------------------------------
def cpu_one_pass(name, kb, host_tables, out):
    """One batch of the workload through the oracle port of the reference's CPU path."""
    import oracle
    F, N, D, B, L, T = WORKLOADS[name]
    salts = salts_for(T)
    for f, (fname, (arena, offs, _)) in enumerate(kb.items()):
        oracle.hashed_bag_forward(arena, offs, B, L, host_tables[f], [N] * T, salts, "sum", mask_empty=True,
                                  out=out, out_col=f * T * D)


def cpu_threads():
    """FIXED OpenMP thread count of the CPU arms: the cores this process may use (affinity / cgroup quota), at most
    32 -- no timing probe, so the reference arm's denominator does not move with a noisy auto-tune."""
    import oracle
    th = oracle.host_threads(cap=32)
    oracle.set_num_threads(th)
    return th


def cpu_reference_run(name, key_batches, host_tables, budget_s=10.0, max_passes=400):
    """Times whole batches of the workload on all host threads; returns (record, last output, its batch index)."""
    import oracle
    F, N, D, B, L, T = WORKLOADS[name]
    out = np.empty((B, F * T * D), dtype=np.float32)
    cpu_one_pass(name, key_batches[0], host_tables, out)          # warm-up (pages the tables in)
    threads = cpu_threads()
    t0, passes = time.perf_counter(), 0
    while passes < max_passes and (passes == 0 or time.perf_counter() - t0 < budget_s):
        cpu_one_pass(name, key_batches[passes % len(key_batches)], host_tables, out)
        passes += 1
    dt = time.perf_counter() - t0
    last = (passes - 1) % len(key_batches)
    oracle.set_num_threads(1)                                     # also the single-thread figure
    t1 = time.perf_counter()
    cpu_one_pass(name, key_batches[last], host_tables, out)
    dt1 = time.perf_counter() - t1
    oracle.set_num_threads(threads)
    return {"value": passes * B / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "single_thread_value": B / dt1, "host_cpus": os.cpu_count(),
            "sample": f"{passes} full batch(es) of {B} samples x {F} fields through oracle/rf_oracle.c "
                      f"(OpenMP, {threads} threads) in {dt:.2f} s (+ 1 batch on 1 thread in {dt1:.2f} s); TensorFlow "
                      f"unavailable, so the reference's TF CPU path is represented by its restated CPU port"}, out, last


def host_tables_numpy(name, seed0=7):
    F, N, D, B, L, T = WORKLOADS[name]
    tabs = []
    for f in range(F):
        row = []
        for t in range(T):
            w = np.random.default_rng(seed0 + f * T + t).random((N, D), dtype=np.float32)
            w -= np.float32(0.5)
            w *= np.float32(0.1)                       # U(-0.05, 0.05), Keras 'uniform'
            row.append(w)
        tabs.append(row)
    return tabs


def run_reference_arm(args):
    """The reference's CPU path (oracle port, fixed thread count) on the repo arm's config.  A step is a bounded
    sample of R whole batches, R sized so that one step takes about 0.25 s: the default 20 steps then time >= 4 s
    of CPU work (round 1 timed 0.87 s and the baseline moved by +-25 % between runs)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    F, N, D, B, L, T = WORKLOADS[name]
    key_batches = make_keys(name, 0)[:2]
    tabs = host_tables_numpy(name)
    out = np.empty((B, F * T * D), dtype=np.float32)
    threads = cpu_threads()
    cpu_one_pass(name, key_batches[0], tabs, out)          # pages the tables in
    t0 = time.perf_counter()
    cpu_one_pass(name, key_batches[1], tabs, out)
    est = max(time.perf_counter() - t0, 1e-4)
    R = int(min(64, max(1, round(0.25 / est))))
    times, n_batches = [], 0
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        for r in range(R):
            cpu_one_pass(name, key_batches[(i * R + r) % len(key_batches)], tabs, out)
        times.append(time.perf_counter() - t0)
        if sum(times) > 150:          # keep the whole run within a few minutes
            break
    timed = times[min(args.warmup, max(len(times) - 1, 0)):]
    ms = 1e3 * float(np.mean(timed))
    value = R * B / (ms / 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(timed), "warmup": min(args.warmup, len(times) - len(timed)), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 pooling / u64 hashing",
            "data": "synthetic", "config": workload_config(name, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"each step = {R} full batch(es) of {B} samples x {F} fields on the host CPU "
                                       f"(oracle/rf_oracle.c, OpenMP, {threads} threads fixed); {len(timed)} steps = "
                                       f"{sum(timed):.2f} s timed; runs on rank 0 only"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# host <-> device copy ceiling of this box, measured with every rank copying at once
# ------------------------------------------------------------------------------------------
def numa_node_of_gpu(index):
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        path = f"/sys/bus/pci/devices/{bus.lower()[-12:]}/numa_node"
        return int(open(path).read())
    except Exception:
        return None


def pcie_diagnostics(dev, world, rank, dist):
    """Plain pinned-memory copies, all ranks at the same time: the D2H / H2D bandwidth each rank gets while the others
    copy too.  It is the ceiling of the e2e number above (436 MB of pooled vectors leave the GPU every step) and shows
    whether the multi-GPU e2e collapse is this box's host side (shared PCIe uplinks / one NUMA node) or ours."""
    import torch
    n = 256 << 20
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    res = {}
    for name, (src, dst) in {"d2h": (d, h), "h2d": (h, d)}.items():
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        gbs = 4 * n / (e0.elapsed_time(e1) / 1e3) / 1e9
        if dist is not None:
            t = torch.tensor([gbs, gbs], dtype=torch.float64, device=dev)
            tmin, tsum = t[:1].clone(), t[1:].clone()
            dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
            dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            res[f"concurrent_{name}_gbs_per_rank_min"] = float(tmin.item())
            res[f"concurrent_{name}_gbs_all_ranks"] = float(tsum.item())
        else:
            res[f"concurrent_{name}_gbs_per_rank_min"] = gbs
            res[f"concurrent_{name}_gbs_all_ranks"] = gbs
    res["gpu_numa_node"] = numa_node_of_gpu(dev.index or 0)
    return res


# ------------------------------------------------------------------------------------------
# C3 end to end (BASELINE.json configs[2]): the full recall-SDPA forward, host keys in, loss out
# ------------------------------------------------------------------------------------------
def run_c3full(dev, steps, warmup, with_cpu=True):
    """base_recall_sdpa two-tower forward at batch 8192 on one GPU: 228 hashed features (2 SipHash tables of 100000 x 8
    each) -> ONE fused bag launch; a hashed behaviour SEQUENCE of up to 50 items -> [B, 50, 64] embeddings ->
    MultiHeadAttention (tcgen05 Dense projections + tcgen05 SDPA) -> mean; towers [1024, 512, 256] (BatchNormalization
    folded, selu, tcgen05 GEMMs, l2 norm in the last epilogue); in-batch softmax loss (tcgen05).  e2e: pinned host key
    buffers -> H2D -> forward -> D2H of the loss, every step.  CPU baseline: the same forward through the oracle."""
    import torch
    from recommendflow_b200 import _native as nat
    from recommendflow_b200.backend.layers.preprocess_layers import HashedEmbeddingBag
    from recommendflow_b200.config_parser import Configuration
    from recommendflow_b200.models.matching.recall_sdpa import RecallSdpa
    from recommendflow_b200.strings import StringColumn
    from recommendflow_b200.synth import PackedBatch, c2_field_keys, decimal_keys

    B, S, dm, NBAT = 8192, 50, 64, 2
    cfg = os.path.join(ROOT, "tests", "golden", "configs", "synth_recall_sdpa")
    conf = Configuration(cfg + ".yaml", slot_map_path=cfg + ".feature.map")
    torch.manual_seed(0)
    model = RecallSdpa(conf, behaviour_dim=dm, num_heads=1)
    names = model.user_cols + model.ad_cols
    beh = HashedEmbeddingBag(100_000, dm, "null", salt=None, mask_value="", mask_zero=True, name="hashing_behaviour")
    model.build(dev)
    beh.build(dev)
    host, devb = [], []
    for bi in range(NBAT):
        rng = np.random.default_rng(555 + bi)
        fields = {}
        for i, n in enumerate(names):
            arena, offs = c2_field_keys(i, B, 1, batch_index=bi)
            fields[n] = (arena, offs, (B, 1))
        lens = rng.integers(1, S + 1, size=B)
        valid = np.arange(S)[None, :] < lens[:, None]
        arena, offs = decimal_keys(b"item_", rng.integers(0, 10**7, size=int(valid.sum())))
        klen = np.zeros(B * S, dtype=np.int64)
        klen[valid.reshape(-1)] = np.diff(offs)
        boffs = np.zeros(B * S + 1, dtype=np.int32)
        boffs[1:] = np.cumsum(klen)                                     # pads are empty strings, like the dataloader's ""
        h = {"keys": PackedBatch.pack(fields, pin=True),
             "beh": StringColumn.from_arena(arena, boffs, (B, S)).pin_memory(),
             "mask": torch.from_numpy(valid.astype(np.float32)[:, :, None].copy()).pin_memory(),
             "y": torch.ones(B).pin_memory(), "np": (fields, arena, boffs, valid)}
        host.append(h)
        devb.append({"keys": h["keys"].to(dev), "beh": h["beh"].to(dev), "mask": h["mask"].to(dev), "y": h["y"].to(dev)})

    def forward(b):
        with torch.no_grad():
            x = beh(b["beh"])                                           # [B, S, 64]: the behaviour sequence's embeddings
            u, a = model.towers(b["keys"], (x, b["mask"]))              # fused bags + SDPA encoder + towers (l2-normalised)
            return model.loss_fun(b["y"], u, a)

    def e2e_step(i):
        h = host[i % NBAT]
        b = {"keys": h["keys"], "beh": h["beh"].to(dev, non_blocking=True), "mask": h["mask"].to(dev, non_blocking=True),
             "y": h["y"].to(dev, non_blocking=True)}
        return float(forward(b).item())                                 # D2H of the scalar + sync: the step's result

    for i in range(max(warmup, 3)):
        forward(devb[i % NBAT])
    torch.cuda.synchronize()
    l0 = nat.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = forward(devb[i % NBAT])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    launches = nat.launch_count() - l0
    for i in range(3):
        e2e_step(i)
    t0 = time.perf_counter()
    for i in range(steps):
        gpu_loss = e2e_step(i)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / steps
    h2d = int(host[0]["keys"].nbytes + host[0]["beh"].nbytes + 4 * host[0]["beh"].offsets.numel() + host[0]["mask"].numel() * 4 + B * 4)

    # ---- the same end to end, PIPELINED: the forward of each rotating batch is recorded once into a CUDA graph over static
    # device buffers (the eager step above is bound by ~1 ms of Python launching 13 kernels); every step the host batch is
    # copied into those buffers on a copy stream (H2D of step i + 1 runs under the forward of step i) and the loss returns
    # through a pinned scalar, which the host reads one step behind.  Every step still moves all its inputs in and its
    # result out inside the timed region.
    piped = None
    try:
        main, sc = torch.cuda.current_stream(dev), torch.cuda.Stream(dev)
        from recommendflow_b200.graphs import GraphedCall
        graphs = [GraphedCall(lambda bi=bi: forward(devb[bi])) for bi in range(NBAT)]      # the public wrapper: record once, replay
        loss_dev = [g.outputs for g in graphs]
        loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(NBAT)]
        copied, done = [None] * NBAT, [None] * NBAT

        def copy_in(i):
            bi = i % NBAT
            h, d = host[bi], devb[bi]
            with torch.cuda.stream(sc):
                if done[bi] is not None:
                    sc.wait_event(done[bi])                              # the forward that last read this buffer set
                h["keys"].to(dev, out=d["keys"])
                d["beh"].data.copy_(h["beh"].data, non_blocking=True)
                d["beh"].offsets.copy_(h["beh"].offsets, non_blocking=True)
                d["mask"].copy_(h["mask"], non_blocking=True)
                d["y"].copy_(h["y"], non_blocking=True)
                copied[bi] = torch.cuda.Event()
                copied[bi].record(sc)

        def launch(i):
            bi = i % NBAT
            main.wait_event(copied[bi])
            graphs[bi]()
            loss_host[bi].copy_(loss_dev[bi], non_blocking=True)
            done[bi] = torch.cuda.Event()
            done[bi].record(main)

        def run(n):
            copy_in(0)
            last = None
            for i in range(n):
                launch(i)
                if i + 1 < n:
                    copy_in(i + 1)                                       # the next batch's H2D runs under this step's forward
                if i > 0:
                    done[(i - 1) % NBAT].synchronize()
                    last = float(loss_host[(i - 1) % NBAT])
            done[(n - 1) % NBAT].synchronize()
            return float(loss_host[(n - 1) % NBAT])

        for bi in range(NBAT):                                           # device-resident: the recorded forward alone
            graphs[bi]()
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for i in range(steps):
            graphs[i % NBAT]()
        g1.record()
        torch.cuda.synchronize()
        graph_ms = g0.elapsed_time(g1) / steps
        run(4)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        piped_loss = run(steps)
        piped_ms = (time.perf_counter() - t0) * 1e3 / steps
        want = float(forward(devb[(steps - 1) % NBAT]).item())
        if piped_loss != want:
            raise RuntimeError(f"pipelined loss {piped_loss} != eager loss {want}")
        piped = {"value": B / (piped_ms / 1e3), "unit": UNIT, "ms_per_step": piped_ms, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                 "device_resident_graph_ms_per_step": graph_ms,
                 "path": "pinned host key arenas (+ mask, labels) -> H2D into static device buffers on a copy stream (prefetch of step "
                         "i + 1 under step i) -> the forward recorded as one CUDA graph per rotating batch -> D2H of the loss scalar "
                         "into pinned memory, read by the host one step behind; loss identical to the eager forward"}
        for g in graphs:
            g.release()
    except Exception as exc:                                             # the synchronous measurement stands on its own
        piped = {"error": f"{type(exc).__name__}: {exc}"}
    res = {"workload": "c3full: base_recall_sdpa two-tower forward, batch 8192: 228 hashed features x 2 tables of 100000 x 8 "
                       "(one fused launch), hashed behaviour sequence <= 50 x 64 -> MultiHeadAttention (tcgen05) -> mean, "
                       "towers [1024, 512, 256] selu + BatchNormalization (tcgen05 Dense, folded), l2 norm, in-batch softmax (tcgen05)",
           "metric": "recall-SDPA forward samples/sec", "unit": UNIT, "value": B / (ms / 1e3), "ms_per_step": ms, "steps": steps,
           "gpu_launches_per_step": launches / steps,
           "e2e": {"value": B / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                   "path": "pinned host key arenas (+ mask, labels) -> H2D -> forward_all / HashedEmbeddingBag / towers / loss -> "
                           "D2H of the loss scalar, one step at a time (host waits for every loss)"}}
    if piped and "error" not in piped:
        # the eager numbers above are bound by the host (13 launches + layer glue from Python per step); the recorded forward
        # is the same computation without that: it becomes the entry's value / e2e, the eager pair stays beside it
        res["eager"] = {"value": res["value"], "ms_per_step": res["ms_per_step"], "e2e": res["e2e"]}
        res["value"], res["ms_per_step"] = B / (piped["device_resident_graph_ms_per_step"] / 1e3), piped["device_resident_graph_ms_per_step"]
        res["e2e"] = piped
    else:
        res["e2e_pipelined"] = piped
    if with_cpu:
        import oracle
        threads = cpu_threads()
        W = {}
        for n in names:
            layer = model.preprocessor[n]
            W[n] = [w for w in layer.get_weights()]
        wb = beh.get_weights()[0]
        att = model.seq_encoder
        proj = [tuple(d.get_weights()) for d in (att.wq, att.wk, att.wv)]

        def tower_stages(seq):
            stages, d = [], None
            for norm, act in seq._stages():
                d = act.dense.kernel.shape[0]
                gamma, beta, mean, var = (t.detach().cpu().numpy() for t in norm.state(d))
                k, bias = act.dense.get_weights()
                stages.append((gamma, beta, mean, var, k, bias, act.activation))
            return stages
        ustages, astages = tower_stages(model.user_dense), tower_stages(model.ad_dense)

        def cpu_forward(h):
            fields, arena, boffs, valid = h["np"]
            cols = []
            for n in names:
                a, o, _ = fields[n]
                layer = model.preprocessor[n]
                cols.append(oracle.hashed_bag_forward(a, o, B, 1, W[n], [layer.num_bins] * 2, list(layer.seeds), layer.combiner))
            ids = oracle.hash_strings(arena, boffs, 100_000, "", None)
            x = oracle.gather_rows(ids, wb).reshape(B, S, dm)
            seq = oracle.multi_head_attention(x, valid.astype(np.float32), *proj, 1).mean(axis=1)
            nu = len(model.user_cols)
            u = oracle.tower_mlp(np.concatenate(cols[:nu] + [seq], axis=1), ustages, eps=1e-6)
            a_ = oracle.tower_mlp(np.concatenate(cols[nu:], axis=1), astages, eps=1e-6)
            return oracle.inbatch_softmax_ce(np.ones(B, np.float32), u, a_, 20.0)[0]

        cpu_loss = cpu_forward(host[(steps - 1) % NBAT])                # also the warm-up
        t0, n = time.perf_counter(), 0
        while n < 8 and (n == 0 or time.perf_counter() - t0 < 12.0):
            cpu_forward(host[n % NBAT])
            n += 1
        dt = time.perf_counter() - t0
        res["cpu_baseline"] = {"value": n * B / dt, "unit": UNIT, "cores": threads, "kind": "port",
                               "sample": f"{n} full batch(es) of {B} samples through the oracle port of the same forward "
                                         f"(oracle/rf_oracle.c: hashing, bags, Dense, SDPA, softmax CE; OpenMP, {threads} threads) "
                                         f"in {dt:.2f} s"}
        res["e2e_vs_cpu_baseline"] = res["e2e"]["value"] / res["cpu_baseline"]["value"]
        rel = abs(gpu_loss - cpu_loss) / max(abs(cpu_loss), 1e-9)
        res["parity_check"] = f"GPU loss {gpu_loss:.6f} vs oracle loss {cpu_loss:.6f} (rel. diff {rel:.2e}; TF32 tensor-core tolerance 2e-2)"
        if not rel <= 2e-2:
            raise SystemExit("bench c3full: " + res["parity_check"])
    del model, beh, devb
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from recommendflow_b200 import _native as nat
    from recommendflow_b200.backend.layers.preprocess_layers import DoubleHashingEmbedding
    from recommendflow_b200.bag_ops import BagPlan, FieldCall
    from recommendflow_b200.synth import PackedBatch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    # NCCL prints its version banner on stdout; the contract is ONE JSON line there, so everything but
    # the final line goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    name = args.workload
    F, N, D, B, L, T = WORKLOADS[name]
    salts = salts_for(T)
    K, W = args.steps, max(args.warmup, 3)

    # ---- resident state: tables (synthetic U(-0.05, 0.05), same values as the CPU arm) --------
    host_tabs = host_tables_numpy(name) if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None
    tables = []
    for f in range(F):
        row = []
        for t in range(T):
            if host_tabs is not None:
                row.append(torch.from_numpy(host_tabs[f][t]).to(dev))
            else:
                g = torch.Generator(device=dev)
                g.manual_seed(7 + f * T + t)
                row.append(torch.empty(N, D, dtype=torch.float32, device=dev).uniform_(-0.05, 0.05, generator=g))
        tables.append(row)

    key_batches = make_keys(name, rank)
    host_packed = [PackedBatch.pack(kb, pin=True) for kb in key_batches]
    dev_packed = [hp.to(dev) for hp in host_packed]
    dev_cols = [dp.columns() for dp in dev_packed]
    out = torch.empty(B, F * T * D, dtype=torch.float32, device=dev)
    names = list(key_batches[0].keys())

    def calls_for(cols):
        return [FieldCall([(tables[f][t], N, salts[t]) for t in range(T)], D, "sum", keys=cols[n],
                          mask_mode=nat.MASK_EMPTY_STRING, out=out[:, f * T * D:(f + 1) * T * D], bag_len=L)
                for f, n in enumerate(names)]

    # descriptors are built once per key batch (the buffers are static); a step is one C call
    all_calls = [BagPlan(calls_for(c), B) for c in dev_cols]

    def step(i):
        all_calls[i % N_KEY_BATCHES].launch()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value, roofline) ------------------------------------------------
    for i in range(W):
        step(i)
    barrier()
    launches0 = nat.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    t_all0, t_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        t_all0.record()
        for i in range(K):
            ev[i][0].record()
            step(W + i)
            ev[i][1].record()
        t_all1.record()
        barrier()
    launches = nat.launch_count() - launches0
    total_ms = t_all0.elapsed_time(t_all1)
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / K
    value = world * B / (ms_per_step / 1e3)

    # ---- end to end THROUGH THE LAYER API: host key buffers in, pooled vectors out to host, every step ---------
    e2e = None
    if not args.no_e2e:
        # The plugin call is `layers.forward_all(batch)` on the mapping get_preprocess_layers returns
        # (backend/utils/preprocess_utils.py): here one HashedEmbeddingBag (T = 1) / DoubleHashingEmbedding (T = 2) per
        # field holding the same tables.  The batch is a pinned host PackedBatch (one arena + one offsets buffer);
        # forward_all copies it to the device (2 copies), builds the launch descriptors and launches the fused kernel.
        # The batch crosses PCIe in E2E_CHUNKS row chunks on two streams so that the D2H of one chunk's pooled vectors
        # overlaps the H2D + kernel of the next (PCIe is full duplex); every step still moves all key bytes in and all
        # pooled vectors out, and the host result is complete at the end of the step.
        from recommendflow_b200.backend.layers.preprocess_layers import HashedEmbeddingBag
        from recommendflow_b200.backend.utils.preprocess_utils import PreprocessLayers
        layers = PreprocessLayers()
        for f, n in enumerate(names):
            if T == 1:
                layer = HashedEmbeddingBag(N, D, "sum", salt=None, mask_value="", mask_zero=True, name=f"hashing_{n}")
                layer.emb.embeddings = torch.nn.Parameter(tables[f][0], requires_grad=False)
            else:
                layer = DoubleHashingEmbedding(num_bins=N, output_dim=D, seeds=[2022, 2023], combiner="sum", mask_value="",
                                               mask_zero=True, name=f"hashing_{n}")
                layer.emb1.embeddings = torch.nn.Parameter(tables[f][0], requires_grad=False)
                layer.emb2.embeddings = torch.nn.Parameter(tables[f][1], requires_grad=False)
            layers[n] = layer
        n_chunks = E2E_CHUNKS if B % E2E_CHUNKS == 0 and B >= 4096 else 1
        rows = B // n_chunks
        host_chunks = []                                   # [batch][chunk] -> pinned PackedBatch of `rows` samples
        for kb in key_batches:
            per = []
            for c in range(n_chunks):
                fields = {}
                for fname, (arena, offs, _) in kb.items():
                    lo, hi = c * rows * L, (c + 1) * rows * L
                    o = offs[lo:hi + 1]
                    fields[fname] = (arena[o[0]:o[-1]], (o - o[0]).astype(np.int32), (rows, L))
                per.append(PackedBatch.pack(fields, pin=True))
            host_chunks.append(per)
        streams = [torch.cuda.Stream(device=dev) for _ in range(E2E_STREAMS)]
        # two pinned result buffers: the caller consumes step i - 1's complete host result while step i is in flight (the
        # input copies of step i then run under the tail of step i - 1's output copies: PCIe is full duplex)
        host_outs = [torch.empty(B, F * T * D, dtype=torch.float32).pin_memory() for _ in range(2)]

        def e2e_step(i):
            bi = i % N_KEY_BATCHES
            host_out = host_outs[i % 2]
            for c, hp in enumerate(host_chunks[bi]):
                with torch.cuda.stream(streams[c % E2E_STREAMS]):
                    layers.forward_all(hp, names=names, out=out[c * rows:(c + 1) * rows])
                    host_out[c * rows:(c + 1) * rows].copy_(out[c * rows:(c + 1) * rows], non_blocking=True)
            done = []
            for st_ in streams:
                e = torch.cuda.Event()
                e.record(st_)
                done.append(e)
            return done

        def consume(done):                    # the host result of that step is complete
            for e in done:
                e.synchronize()

        torch.cuda.synchronize()
        for i in range(3):
            consume(e2e_step(i))
        barrier()
        ke = max(3, min(K, 20))
        launches_e2e0 = nat.launch_count()
        t0 = time.perf_counter()
        prev = None
        for i in range(ke):
            cur_done = e2e_step(i)
            if prev is not None:
                consume(prev)                 # every step's result reaches the host inside the timed region, one step behind
            prev = cur_done
        consume(prev)
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / ke
        # the layer call's host result equals the kernel-only path's, bit for bit
        host_out = host_outs[(ke - 1) % 2]
        step((ke - 1) % N_KEY_BATCHES)
        torch.cuda.synchronize()
        if not torch.equal(host_out, out.cpu()):
            raise SystemExit("bench: e2e (forward_all) output differs from the BagPlan launch")
        if world > 1:
            t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
        e2e = {"value": world * B / (e2e_ms / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": int(np.mean([sum(h.nbytes for h in per) for per in host_chunks])),
               "d2h_bytes_per_step": int(host_out.numel() * 4), "ms_per_step": e2e_ms, "steps": ke,
               "gpu_launches": int(nat.launch_count() - launches_e2e0),
               "path": f"pinned host PackedBatch -> PreprocessLayers.forward_all (H2D of arena + offsets, descriptors, "
                       f"rf_bag_forward) -> D2H of the pooled [B, sum(T*D)] fp32; {n_chunks} row chunks on {E2E_STREAMS} streams "
                       f"(copy/compute overlap), two pinned result buffers: the host consumes step i - 1's complete result while "
                       f"step i is in flight"}
        e2e.update(pcie_diagnostics(dev, world, rank, dist if world > 1 else None))
        # the step cannot beat the slowest rank's share of the box's host links: all pooled vectors out + all keys in
        floor_ms = (e2e["d2h_bytes_per_step"] / (e2e["concurrent_d2h_gbs_per_rank_min"] * 1e6)
                    if e2e["concurrent_d2h_gbs_per_rank_min"] > 0 else None)
        e2e["box_d2h_floor_ms_per_step"] = floor_ms
        e2e["fraction_of_box_d2h_ceiling"] = None if not floor_ms else floor_ms / e2e_ms
        e2e["note"] = ("bound by the D2H of the pooled vectors (436 MB per rank per step): the plain-copy bandwidth measured with "
                       "all ranks copying at once is printed beside it (one GPU: ~57 GB/s = PCIe Gen5 x16; eight GPUs of this box "
                       "share ~123 GB/s of D2H)")
        del layers

    # ---- roofline of the one kernel ------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    bps = algorithmic_bytes_per_sample(name, key_batches)
    achieved = bps * B / (kern_ms / 1e3) / 1e9
    roofline = {"bound": "hbm", "kernel": "rf::bag_forward_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s",
                "frac_of_nominal_8TBs": achieved / 8000.0, "algorithmic_bytes_per_sample": bps,
                "algorithmic_bytes_per_launch": bps * B, "kernel_ms": kern_ms, "traffic": None}
    if name in KEY_ZIPF:
        roofline["note"] = ("Zipf keys: the hot rows are served by L2, so the algorithmic bytes (every gathered row "
                            "counted) exceed the HBM traffic and frac may exceed 1; the uniform-key workload c2 is the "
                            "HBM-bound measurement")
    try:      # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu --set full capture
        tr = json.load(open(os.path.join(ROOT, "profiles", f"traffic_{name}.json")))
        roofline["traffic"] = tr["traffic_bytes_per_launch"]
        roofline["traffic_source"] = tr["source"]
    except Exception:
        pass

    # ---- CPU baseline beside it (rank 0, N == 1 only) ------------------------------------------
    cpu, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, cpu_out, bi = cpu_reference_run(name, key_batches, host_tabs)
        # the checker comes for free: the GPU's result for the same batch must equal the CPU arm's
        step(bi)
        torch.cuda.synchronize()
        same = bool(np.array_equal(out.cpu().numpy().view(np.uint32), cpu_out.view(np.uint32)))
        parity = f"GPU output {'==' if same else '!='} CPU arm output, bit for bit, on key batch {bi} ({B} x {F * T * D} fp32)"
        if not same:
            raise SystemExit("bench: " + parity)

    # ---- C4: jagged bags, mean pooling, 100M x 128 table: one GPU = the fused kernel on
    # the whole table; N > 1 = the table row-sharded id % N with the p2p (NVLink peer memory) exchange.
    c4 = None
    if not args.no_c4 and name == "c2":
        e2e_plans = None
        del tables, all_calls, dev_cols, dev_packed, out
        torch.cuda.empty_cache()
        from tools.bench_sharded import parse as c4_parse, run as c4_run
        c4 = c4_run(c4_parse(["--steps", "20", "--warmup", "5"]), world, rank, dev)

    # ---- C3 end to end (host keys -> bags -> SDPA -> towers -> loss -> host), rank 0 of a single-GPU run ----
    c3 = None
    if not args.no_c3 and name == "c2" and world == 1:
        c3 = run_c3full(dev, int(os.environ.get("RF_BENCH_C3_STEPS", "20")), int(os.environ.get("RF_BENCH_C3_WARMUP", "5")),
                        with_cpu=not args.no_cpu_baseline)

    # ---- C3 training step (forward + backward + Keras Adam on every variable), eager and as one CUDA graph ----
    train = None
    if not args.no_train and name == "c2" and world == 1:
        torch.cuda.empty_cache()
        from tools.bench_train import run as train_run
        train = train_run(steps=10)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 pooling / u64 hashing", "data": "synthetic", "config": workload_config(name, world),
                "gpu_launches": int(launches), "clocks": clocks.summary(), "e2e": e2e, "roofline": roofline,
                "cpu_baseline": cpu, "parity_check": parity, "sharded_c4": c4, "c3full": c3, "train_c3": train}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
