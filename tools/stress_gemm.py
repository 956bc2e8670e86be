#!/usr/bin/env python
"""Randomised shapes through the tcgen05 Dense kernel (single CTA, CTA pairs, split-K) against a float64 product.
    python tools/stress_gemm.py [--n 150] [--seed 0]          RF_DENSE_PAIR=2 forces pairs wherever legal"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=150)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    import numpy as np
    import torch
    from recommendflow_b200.dense_ops import dense_forward
    rng = np.random.default_rng(args.seed)
    acts = [None, "relu", "selu", "tanh", "sigmoid"]
    worst, bad = 0.0, []
    for it in range(args.n):
        M = int(rng.choice([rng.integers(1, 300), rng.integers(300, 3000), rng.integers(3000, 20000)]))
        K = int(rng.integers(1, 600)) * 4
        N = int(rng.integers(1, 300)) * 4
        act = acts[int(rng.integers(0, len(acts)))]
        with_bias = bool(rng.integers(0, 2))
        l2 = bool(rng.integers(0, 4) == 0) and N <= 256
        x = torch.randn(M, K, device="cuda")
        wt = torch.randn(N, K, device="cuda") / K ** 0.5
        b = torch.randn(N, device="cuda") * 0.1 if with_bias else None
        got = dense_forward(x, wt, b, act, l2_normalize=l2).double()
        z = x.double() @ wt.double().t() + (b.double() if b is not None else 0)
        want = {None: lambda t: t, "relu": torch.relu, "selu": torch.selu, "tanh": torch.tanh, "sigmoid": torch.sigmoid}[act](z)
        if l2:
            want = want / want.norm(dim=1, keepdim=True).clamp_min(1e-12)
        err = float((got - want).abs().max())
        tol = 2e-2 * max(1.0, float(want.abs().max())) * (1.0 if not l2 else 0.2)
        worst = max(worst, err / tol)
        if not err <= tol or not bool(torch.isfinite(got).all()):
            bad.append({"M": M, "K": K, "N": N, "act": act, "bias": with_bias, "l2": l2, "err": err, "tol": tol})
    print(json.dumps({"shapes": args.n, "pair_env": os.environ.get("RF_DENSE_PAIR"), "worst_err_over_tol": worst, "failures": bad}))
    if bad:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
