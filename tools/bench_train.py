#!/usr/bin/env python
"""Device timing of the backward kernels and of one full training step of the recall-SDPA model at the C3
shape (batch 8192, 228 hashed features x 2 tables of 100000 x 8, [B, 50, 64] behaviour sequence, towers
[1024, 512, 256]) on one B200.  Prints one JSON line.  STEPS / LAZY env vars."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, steps, warmup=3):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def run(steps=10, lazy=False, graph=True, profile_path=None):
    """The measurements as a dict (bench.py appends it to its line as `train_c3`)."""
    import torch
    from recommendflow_b200 import _native as nat
    from recommendflow_b200.config_parser import Configuration
    from recommendflow_b200.dense_ops import inbatch_rowstats, inbatch_softmax_ce_backward, sdpa_backward
    from recommendflow_b200.models.matching.recall_sdpa import RecallSdpa
    from recommendflow_b200.strings import StringColumn
    from recommendflow_b200.synth import c2_field_keys
    from recommendflow_b200.training import RecallSdpaTrainer

    B, S, dm = 8192, 50, 64
    out = {"workload": "c3 training step: base_recall_sdpa two-tower model, batch 8192, 228 hashed features x 2 tables of 100000 x 8, "
                       "SDPA encoder [B,50,64], towers [1024,512,256] with BatchNormalization (batch statistics) + dropout 0.3, in-batch "
                       "softmax loss, Keras Adam on every variable (train.py:97-104)", "batch": B, "steps": steps}
    # kernels alone
    q = torch.nn.functional.normalize(torch.randn(B, 256, device="cuda"), dim=1)
    d = torch.nn.functional.normalize(torch.randn(B, 256, device="cuda"), dim=1)
    y = torch.ones(B, device="cuda")
    lse = inbatch_rowstats(q, d, y_true=y, want=("lse",))["lse"]
    out["ce_backward_ms"] = timed(lambda: inbatch_softmax_ce_backward(q, d, y, lse), steps)
    out["ce_backward_tflops"] = 4 * 2 * B * B * 256 / (out["ce_backward_ms"] / 1e3) / 1e12      # S twice + two products
    x = torch.randn(B, S, dm, device="cuda")
    g = torch.randn(B, S, dm, device="cuda")
    mask = (torch.arange(S, device="cuda")[None, :, None] < torch.randint(1, S + 1, (B, 1, 1), device="cuda")).float()
    out["sdpa_backward_ms"] = timed(lambda: sdpa_backward(x, x, x, mask, g, precision="tf32"), steps)
    out["sdpa_backward_fp32_ms"] = timed(lambda: sdpa_backward(x, x, x, mask, g, precision="fp32"), steps)
    # the whole step
    cfg = os.path.join(ROOT, "tests", "golden", "configs", "synth_recall_sdpa")
    conf = Configuration(cfg + ".yaml", slot_map_path=cfg + ".feature.map")
    model = RecallSdpa(conf, behaviour_dim=dm, num_heads=1)
    trainer = RecallSdpaTrainer(model, learning_rate=1e-4, lazy_embedding_adam=lazy)
    names = model.user_cols + model.ad_cols
    batch = {}
    for i, n in enumerate(names):
        arena, offs = c2_field_keys(i, B, 1)
        batch[n] = StringColumn.from_arena(arena, offs, (B, 1)).to("cuda")
    l0 = None

    def step():
        return trainer.train_step(batch, y, (x, mask))
    step()
    l0 = nat.launch_count()
    out["train_step_ms"] = timed(step, steps, warmup=2)
    out["train_samples_per_s"] = B / (out["train_step_ms"] / 1e3)
    out["launches_per_step"] = (nat.launch_count() - l0) / (steps + 2)
    out["embedding_adam"] = "lazy" if trainer.lazy else "keras (all rows decay)"
    out["loss_after"] = float(step())
    if graph:        # the same step recorded into one CUDA graph (training.GraphedTrainStep)
        from recommendflow_b200.training import GraphedTrainStep
        graphed = GraphedTrainStep(trainer, batch, y, (x, mask), warmup=2)
        out["train_step_graphed_ms"] = timed(graphed, steps, warmup=2)
        out["train_graphed_samples_per_s"] = B / (out["train_step_graphed_ms"] / 1e3)
        out["loss_after_graphed"] = float(graphed())
        step = graphed
    if profile_path:          # where the step's time goes (kineto; not a timing source for the numbers above)
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                step()
            torch.cuda.synchronize()
        with open(profile_path, "w") as f:
            f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
            f.write("\n\n")
            f.write(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=30, max_name_column_width=70))

    return out


def main():
    print(json.dumps(run(int(os.environ.get("STEPS", "10")), os.environ.get("LAZY", "0") == "1", os.environ.get("GRAPH", "1") == "1",
                         os.environ.get("PROFILE"))))


if __name__ == "__main__":
    main()
