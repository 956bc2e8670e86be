#!/usr/bin/env python
"""cProfile of the EAGER C3 forward (host side): where the ~1.1 ms per step of Python goes.  python tools/profile_eager.py"""
import cProfile
import io
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from recommendflow_b200.config_parser import Configuration
    from recommendflow_b200.models.matching.recall_sdpa import RecallSdpa
    from recommendflow_b200.synth import PackedBatch, c2_field_keys
    B, S, dm = 8192, 50, 64
    dev = torch.device("cuda", 0)
    cfg = os.path.join(ROOT, "tests", "golden", "configs", "synth_recall_sdpa")
    conf = Configuration(cfg + ".yaml", slot_map_path=cfg + ".feature.map")
    torch.manual_seed(0)
    model = RecallSdpa(conf, behaviour_dim=dm, num_heads=1)
    names = model.user_cols + model.ad_cols
    model.build(dev)
    fields = {}
    for i, n in enumerate(names):
        arena, offs = c2_field_keys(i, B, 1)
        fields[n] = (arena, offs, (B, 1))
    keys = PackedBatch.pack(fields).to(dev)
    x = torch.randn(B, S, dm, device=dev)
    mask = torch.ones(B, S, 1, device=dev)
    y = torch.ones(B, device=dev)

    def forward():
        with torch.no_grad():
            u, a = model.towers(keys, (x, mask))
            return model.loss_fun(y, u, a)

    for _ in range(20):
        forward()
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(300):
        forward()
    pr.disable()
    torch.cuda.synchronize()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(35)
    print(s.getvalue())


if __name__ == "__main__":
    main()
