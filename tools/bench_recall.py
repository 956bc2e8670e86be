#!/usr/bin/env python
"""C3 (BASELINE.json configs[2]): the recall-SDPA two-tower forward at batch 8192 on one B200:
228 hashed features (normalised base_recall_sdpa plan) -> fused bags -> SDPA encoder over a
[B, 50, 64] behaviour sequence -> tower MLPs [1024, 512, 256] -> l2 norm -> in-batch softmax loss.
Prints one JSON line with per-stage device times (CUDA events) and samples/s."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from recommendflow_b200 import _native as nat
    from recommendflow_b200.bag_ops import BagPlan
    from recommendflow_b200.config_parser import Configuration
    from recommendflow_b200.models.matching.recall_sdpa import RecallSdpa
    from recommendflow_b200.strings import StringColumn
    from recommendflow_b200.synth import c2_field_keys

    B, S, dm, steps = 8192, 50, 64, int(os.environ.get("STEPS", "20"))
    cfg = os.path.join(ROOT, "tests", "golden", "configs", "synth_recall_sdpa")
    conf = Configuration(cfg + ".yaml", slot_map_path=cfg + ".feature.map")
    model = RecallSdpa(conf, behaviour_dim=dm, num_heads=1)
    names = model.user_cols + model.ad_cols
    batch = {}
    for i, n in enumerate(names):
        arena, offs = c2_field_keys(i, B, 1)
        batch[n] = StringColumn.from_arena(arena, offs, (B, 1)).to("cuda")
    x = torch.randn(B, S, dm, device="cuda")
    mask = (torch.arange(S, device="cuda")[None, :, None] < torch.randint(1, S + 1, (B, 1, 1), device="cuda")).float()
    y = torch.ones(B, device="cuda")
    layers = model.preprocessor
    layout, total = layers.output_layout(names)
    fused = torch.empty(B, total, device="cuda")
    for n in names:
        layers[n].build(torch.device("cuda"))
    plan = BagPlan([layers[n].field_call(batch[n], fused[:, layout[n][0]:layout[n][0] + layout[n][1]]) for n in names], B)
    ucols = sum(layout[n][1] for n in model.user_cols)

    def stage_bags():
        plan.launch()

    def stage_sdpa():
        return model.seq_encoder(x, x, x, mask).mean(dim=1)

    def stage_towers(seq):
        u = torch.cat([fused[:, :ucols], seq], dim=-1)
        return model.embedding_norm(model.user_dense(u)), model.embedding_norm(model.ad_dense(fused[:, ucols:]))

    def stage_loss(u, a):
        return model.loss_fun(y, u, a)

    def step(ev=None):
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if ev is not None else None
        if marks: marks[0].record()
        stage_bags()
        if marks: marks[1].record()
        seq = stage_sdpa()
        if marks: marks[2].record()
        u, a = stage_towers(seq)
        if marks: marks[3].record()
        loss = stage_loss(u, a)
        if marks:
            marks[4].record()
            ev.append(marks)
        return loss

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    l0 = nat.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    launches = nat.launch_count() - l0
    ev = []
    for _ in range(5):
        step(ev)
    torch.cuda.synchronize()
    stages = ["bags_228_fields", "sdpa_encoder", "tower_mlps_cublas", "inbatch_softmax_loss"]
    per = {s: float(np.mean([m[i].elapsed_time(m[i + 1]) for m in ev])) for i, s in enumerate(stages)}
    print(json.dumps({"workload": "c3: recall-SDPA two-tower forward, batch 8192, 228 hashed features x 2 tables of 100000 x 8, "
                                  "SDPA encoder [B,50,64], towers [1024,512,256], in-batch softmax (tf32 tensor cores)",
                      "ms_per_step": ms, "samples_per_s": B / (ms / 1e3), "stage_ms": per, "loss": float(loss),
                      "gpu_launches_per_step": launches / steps}))


if __name__ == "__main__":
    main()
