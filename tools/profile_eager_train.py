#!/usr/bin/env python
"""cProfile of the EAGER C3 training step (host side).  python tools/profile_eager_train.py"""
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
    from recommendflow_b200.strings import StringColumn
    from recommendflow_b200.synth import c2_field_keys
    from recommendflow_b200.training import RecallSdpaTrainer
    B, S, dm = 8192, 50, 64
    cfg = os.path.join(ROOT, "tests", "golden", "configs", "synth_recall_sdpa")
    conf = Configuration(cfg + ".yaml", slot_map_path=cfg + ".feature.map")
    model = RecallSdpa(conf, behaviour_dim=dm, num_heads=1)
    trainer = RecallSdpaTrainer(model, learning_rate=1e-4)
    names = model.user_cols + model.ad_cols
    batch = {}
    for i, n in enumerate(names):
        arena, offs = c2_field_keys(i, B, 1)
        batch[n] = StringColumn.from_arena(arena, offs, (B, 1)).to("cuda")
    x = torch.randn(B, S, dm, device="cuda")
    mask = torch.ones(B, S, 1, device="cuda")
    y = torch.ones(B, device="cuda")
    for _ in range(5):
        trainer.train_step(batch, y, (x, mask))
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(30):
        trainer.train_step(batch, y, (x, mask))
    pr.disable()
    torch.cuda.synchronize()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(40)
    print(s.getvalue())


if __name__ == "__main__":
    main()
