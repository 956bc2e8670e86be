#!/usr/bin/env python
"""C3's dense rows on one GPU: the in-batch softmax loss (B x B logits on tcgen05) and SDPA.

    python tools/bench_logits.py [--batch 8192] [--dim 256] [--steps 20]
Prints one JSON line with ms, TFLOP/s and the fraction of the measured bf16/TF32-class peak."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--dim", type=int, default=256)
    ap.add_argument("--seq", type=int, default=50)
    ap.add_argument("--d-model", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--skip-fp32", action="store_true")
    ap.add_argument("--big", action="store_true", help="also time B = 65536")
    args = ap.parse_args()
    import torch
    from recommendflow_b200 import _native as nat
    from recommendflow_b200.dense_ops import inbatch_rowstats, sdpa

    torch.manual_seed(0)
    B, D = args.batch, args.dim
    q = torch.nn.functional.normalize(torch.randn(B, D, device="cuda"), dim=1)
    d = torch.nn.functional.normalize(0.7 * q + 0.7 * torch.randn(B, D, device="cuda"), dim=1)
    y = torch.ones(B, device="cuda")

    def timed(fn, n):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, r

    out = {"workload": f"in-batch softmax loss, B={B}, Dt={D}, scale 20 (match_losses.py:150-165)"}
    flops = 2.0 * B * B * D
    for prec in (["tf32", "bf16"] if args.skip_fp32 else ["tf32", "bf16", "fp32"]):
        ms, r = timed(lambda: inbatch_rowstats(q, d, y_true=y, precision=prec), args.steps)
        out[prec] = {"ms": ms, "tflops": flops / ms / 1e9, "loss": float(r["loss"])}
    if args.big:
        Bb = 65536
        qb = torch.nn.functional.normalize(torch.randn(Bb, D, device="cuda"), dim=1)
        db = torch.nn.functional.normalize(0.7 * qb + 0.7 * torch.randn(Bb, D, device="cuda"), dim=1)
        yb = torch.ones(Bb, device="cuda")
        for prec in ("tf32", "bf16"):
            ms, r = timed(lambda: inbatch_rowstats(qb, db, y_true=yb, precision=prec), 5)
            out[f"{prec}_B65536"] = {"ms": ms, "tflops": 2.0 * Bb * Bb * D / ms / 1e9, "loss": float(r["loss"])}
        del qb, db
    S, dm = args.seq, args.d_model
    qq, kk, vv = (torch.randn(B, S, dm, device="cuda") for _ in range(3))
    mask = (torch.arange(S, device="cuda")[None, :, None] < torch.randint(1, S + 1, (B, 1, 1), device="cuda")).float()
    runs = [timed(lambda: sdpa(qq, kk, vv, mask), args.steps)[0] for _ in range(5)]      # five back-to-back measurements
    ms = min(runs)
    out["sdpa"] = {"shape": [B, S, dm], "ms": ms, "runs_ms": [round(r, 4) for r in runs], "gflop": 4.0 * B * S * S * dm / 1e9,
                   "hbm_gbs": 4 * B * S * dm * 4 / ms / 1e6}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        out["tf32"]["frac_of_measured_bf16_peak"] = out["tf32"]["tflops"] / peaks["bf16_tflops"]
        out["bf16"]["frac_of_measured_bf16_peak"] = out["bf16"]["tflops"] / peaks["bf16_tflops"]
    except Exception:
        pass
    out["gpu_launches"] = nat.launch_count()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
