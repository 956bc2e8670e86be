"""Device timing of the bag backward + Adam update (rf_bag_backward_adam) on one table:
python tools/bench_adam.py [--rows N --dim D --batch B --bag-len L --steps K]
Prints one JSON line: ms per step for Keras (dense) and lazy semantics, and the HBM rate of the dense pass."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from recommendflow_b200 import _native as nat
    from recommendflow_b200.bag_ops import BagAdam
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=64)
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--bag-len", type=int, default=4)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    table = torch.empty(a.rows, a.dim, device=dev).uniform_(-0.05, 0.05)
    gen = torch.Generator(device=dev).manual_seed(1)
    ids = [torch.randint(1, a.rows, (a.batch * a.bag_len,), device=dev, generator=gen) for _ in range(4)]
    grad = torch.randn(a.batch, a.dim, device=dev, generator=gen)
    out = {"rows": a.rows, "dim": a.dim, "batch": a.batch, "bag_len": a.bag_len, "steps": a.steps}
    for lazy in (False, True):
        opt = BagAdam(table, learning_rate=1e-4, lazy=lazy)
        for i in range(3):
            opt.apply(ids[i % 4], grad, "sum", bag_len=a.bag_len)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.steps):
            opt.apply(ids[i % 4], grad, "sum", bag_len=a.bag_len)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        out["lazy_ms" if lazy else "keras_ms"] = ms
        if not lazy:
            out["keras_table_gbs"] = 6 * 4 * a.rows * a.dim / (ms / 1e3) / 1e9      # read + write of w, m, v
    out["gpu_launches"] = nat.launch_count()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
