#!/usr/bin/env python
"""The tcgen05 Dense kernel (rf_dense_forward_tc) on the C3 tower / projection shapes: ms, TFLOP/s, GB/s per shape.

    python tools/bench_gemm.py [--steps 20]
Prints one JSON line."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    import torch
    from recommendflow_b200.dense_ops import dense_forward
    ap.add_argument("--train", action="store_true", help="the products of one C3 training step (forward, dX, dW, CE backward)")
    args = ap.parse_args()
    shapes = [(8192, 1888, 1024, "selu"), (8192, 1024, 512, "selu"), (8192, 512, 256, "selu"), (409600, 64, 64, None),
              (65536, 1888, 1024, "selu")]
    if args.train:      # (rows, contraction, columns): dX = dZ W^T, dW = X^T dZ (contraction over the batch), q|k|v projection, CE slab products
        shapes = [(8192, 1024, 1888, None), (8192, 512, 1024, None), (8192, 256, 512, None),
                  (1888, 8192, 1024, None), (1024, 8192, 512, None), (512, 8192, 256, None),
                  (409600, 64, 192, "bias"), (409600, 192, 64, None), (64, 409600, 192, None),
                  (8192, 256, 8192, None), (8192, 8192, 256, None)]
    out = {}
    for M, K, N, act in shapes:
        x = torch.randn(M, K, device="cuda")
        w = torch.randn(N, K, device="cuda") * 0.02
        b = torch.zeros(N, device="cuda") if act is not None else None
        act = None if act == "bias" else act
        y = torch.empty(M, N, device="cuda")
        for _ in range(3):
            dense_forward(x, w, b, act, out=y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            dense_forward(x, w, b, act, out=y)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        wt = w.t().contiguous()
        for _ in range(3):
            torch.matmul(x, wt)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            torch.matmul(x, wt)
        e1.record()
        torch.cuda.synchronize()
        torch.backends.cuda.matmul.allow_tf32 = prev
        lib_ms = e0.elapsed_time(e1) / args.steps
        out[f"{M}x{K}x{N}"] = {"ms": ms, "tflops": 2.0 * M * K * N / ms / 1e9, "gbs": 4.0 * (M * K + N * K + M * N) / ms / 1e6,
                               "cublas_tf32_matmul_only_ms": lib_ms}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
