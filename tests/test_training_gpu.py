"""A few optimisation steps of the two-tower recall model (recommendflow_b200/training.py): CUDA forward,
CUDA backward of the loss / SDPA / bags, Keras-Adam updates.  The task is learnable by construction (the ad's
item id determines the user's clicked item), so the in-batch softmax loss must fall well below ln(B)."""
import os

import numpy as np
import pytest
import torch

from recommendflow_b200.config_parser import Configuration
from recommendflow_b200.models.matching.recall_sdpa import RecallSdpa
from recommendflow_b200.strings import StringColumn
from recommendflow_b200.training import RecallSdpaTrainer

pytestmark = pytest.mark.gpu


def test_recall_sdpa_training_steps_reduce_the_loss(golden_dir):
    torch.manual_seed(3)
    rng = np.random.default_rng(3)
    conf = Configuration(os.path.join(golden_dir, "configs", "synth_recall_sdpa.yaml"),
                         slot_map_path=os.path.join(golden_dir, "configs", "synth_recall_sdpa.feature.map"))
    keep = set(conf.features.user_feature_names[:3] + conf.features.ad_feature_names[:2])
    for f in conf.features.features:
        if f.is_hashing() and f.name not in keep:
            f.working = False
    B, S, dm = 256, 10, 32
    model = RecallSdpa(conf, tower_units=(64, 32), behaviour_dim=dm, num_heads=2)
    trainer = RecallSdpaTrainer(model, learning_rate=5e-3)
    n_items = 400

    def make_batch():
        item = rng.integers(0, n_items, size=B)
        batch = {}
        for i, name in enumerate(model.user_cols):           # user features: noisy functions of the item
            batch[name] = StringColumn.from_lists([[f"u{i}_{v}", f"u{i}_{v % 7}"] for v in item]).to("cuda")
        for i, name in enumerate(model.ad_cols):
            batch[name] = StringColumn.from_lists([[f"a{i}_{v}"] for v in item]).to("cuda")
        x = torch.from_numpy(rng.standard_normal((B, S, dm)).astype(np.float32)).cuda()
        mask = torch.from_numpy((np.arange(S)[None, :, None] < rng.integers(1, S + 1, size=(B, 1, 1))).astype(np.float32)).cuda()
        return batch, torch.ones(B, device="cuda"), (x, mask)

    losses = []
    for step in range(60):
        batch, y, behaviour = make_batch()
        losses.append(float(trainer.train_step(batch, y, behaviour)))
    assert np.isfinite(losses).all()
    head, tail = np.mean(losses[:5]), np.mean(losses[-5:])
    # duplicates of an item inside a batch are indistinguishable positives, so the floor is above zero
    assert head > 4.0 and tail < head - 1.0, (head, tail)
    assert trainer.iterations == 60 and len(trainer.bag_opts) == 1          # one group: all ten tables are 8 wide
    group, members = next(iter(trainer.bag_opts.values()))
    assert len(members) == 2 * 5 and group.iterations == 60 and all(float(v.abs().sum()) > 0 for v in group.v)
    # inference mode is restored after every step
    assert all(not m.batch_stats for m in trainer._modules(type(model.user_dense.layers[0])))
    out = model(batch, y_true=y, behaviour=behaviour, training=False)
    assert out["user"].shape == (B, 32)


def test_checkpoint_round_trip_restores_weights_and_optimizer_state(golden_dir, tmp_path):
    """model.state_dict() (tables, tower kernels, BatchNormalization parameters + moving statistics) and
    trainer.state_dict() (Adam moments of dense variables and of every table, iteration counters) saved after 3 steps
    and loaded into a FRESH model / trainer: the next step's loss and the weights after it are identical, bit for bit
    (dropout off so that the two runs see the same arithmetic)."""
    conf_path = os.path.join(golden_dir, "configs", "synth_recall_sdpa.yaml")
    map_path = os.path.join(golden_dir, "configs", "synth_recall_sdpa.feature.map")

    def fresh(seed):
        torch.manual_seed(seed)
        conf = Configuration(conf_path, slot_map_path=map_path)
        keep = set(conf.features.user_feature_names[:2] + conf.features.ad_feature_names[:2])
        for f in conf.features.features:
            if f.is_hashing() and f.name not in keep:
                f.working = False
        model = RecallSdpa(conf, tower_units=(32, 16))
        for m in model.modules():
            if hasattr(m, "rate"):
                m.rate = 0.0
        return model, RecallSdpaTrainer(model, learning_rate=1e-2)

    rng = np.random.default_rng(9)
    B = 128

    def make_batch(model):
        item = rng.integers(0, 50, size=B)
        batch = {n: StringColumn.from_lists([[f"{n}_{v}"] for v in item]).to("cuda") for n in model.user_cols + model.ad_cols}
        return batch, torch.ones(B, device="cuda")

    model, trainer = fresh(1)
    batches = [make_batch(model) for _ in range(5)]
    for b, y in batches[:3]:
        trainer.train_step(b, y)
    path = tmp_path / "ckpt.pt"
    torch.save({"model": model.state_dict(), "trainer": trainer.state_dict()}, path)
    keys = set(model.state_dict().keys())
    assert any("embeddings" in k for k in keys) and any("moving_mean" in k for k in keys) and any("kernel" in k for k in keys)
    want_loss = [float(trainer.train_step(b, y)) for b, y in batches[3:]]
    want = {k: v.clone() for k, v in model.state_dict().items()}

    model2, trainer2 = fresh(2)                      # different initialisation: everything must come from the file
    b0, y0 = batches[0]
    with torch.no_grad():
        model2(b0, y_true=y0, training=True)         # builds the lazily created variables
    ckpt = torch.load(path)
    model2.load_state_dict(ckpt["model"])
    trainer2.load_state_dict(ckpt["trainer"], example_batch=(b0, y0, None))
    got_loss = [float(trainer2.train_step(b, y)) for b, y in batches[3:]]
    assert got_loss == want_loss
    got = model2.state_dict()
    assert set(got) == set(want)
    for k in want:
        assert torch.equal(got[k], want[k]), k


def test_graphed_train_step_replays_the_eager_step(golden_dir):
    """training.GraphedTrainStep: the whole step (fused bag forward, towers, loss, autograd, both Adam updates with their
    step counters on the device) recorded into ONE CUDA graph; replays on refilled static inputs follow the eager
    trainer's losses and end on the same weights."""
    from recommendflow_b200 import _native as nat
    from recommendflow_b200.training import GraphedTrainStep
    conf_path = os.path.join(golden_dir, "configs", "synth_recall_sdpa.yaml")
    map_path = os.path.join(golden_dir, "configs", "synth_recall_sdpa.feature.map")

    def fresh():
        torch.manual_seed(4)
        conf = Configuration(conf_path, slot_map_path=map_path)
        keep = set(conf.features.user_feature_names[:3] + conf.features.ad_feature_names[:2])
        for f in conf.features.features:
            if f.is_hashing() and f.name not in keep:
                f.working = False
        model = RecallSdpa(conf, tower_units=(64, 32), behaviour_dim=32, num_heads=1)
        for m in model.modules():
            if hasattr(m, "rate"):
                m.rate = 0.0
        return model, RecallSdpaTrainer(model, learning_rate=5e-3)

    rng = np.random.default_rng(21)
    B, S, dm = 512, 10, 32

    def make_batch(model):
        item = rng.integers(0, 300, size=B)
        batch = {n: StringColumn.from_lists([[f"{n[:4]}_{v:04d}"] for v in item]).to("cuda") for n in model.user_cols + model.ad_cols}
        x = torch.from_numpy(rng.standard_normal((B, S, dm)).astype(np.float32)).cuda()
        mask = torch.from_numpy((np.arange(S)[None, :, None] < rng.integers(1, S + 1, size=(B, 1, 1))).astype(np.float32)).cuda()
        return batch, torch.ones(B, device="cuda"), (x, mask)

    eager_model, eager = fresh()
    batches = [make_batch(eager_model) for _ in range(5)]
    b0, y0, beh0 = batches[0]
    want = [float(eager.train_step(b0, y0, beh0)) for _ in range(3)]           # the graphed trainer's first step + 2 warm-up steps
    want += [float(eager.train_step(b, y, beh)) for b, y, beh in batches[1:]]

    model, trainer = fresh()
    static = ({n: StringColumn(c.data.clone(), c.offsets.clone(), c.shape) for n, c in b0.items()}, y0.clone(),
              (beh0[0].clone(), beh0[1].clone()))
    step = GraphedTrainStep(trainer, *static, warmup=2)
    assert trainer.iterations == 3
    got = []
    before = nat.launch_count()
    for b, y, beh in batches[1:]:
        for n, c in b.items():                                                 # refill the static inputs in place
            static[0][n].data.copy_(c.data)
            static[0][n].offsets.copy_(c.offsets)
        static[1].copy_(y)
        static[2][0].copy_(beh[0])
        static[2][1].copy_(beh[1])
        got.append(float(step()))
    assert nat.launch_count() == before, "a replay goes through no host-side launch"
    assert trainer.iterations == 3 + len(got) == eager.iterations
    # the device-side step computes lr_t in float64 tensor ops and applies it as (m / den) * lr_t instead of addcdiv: last-bit
    # differences per step, which Adam's sign-like update turns into a fraction of lr on the weights whose gradient is noise
    np.testing.assert_allclose(got, want[3:], rtol=3e-3)
    a, b = eager_model.state_dict(), model.state_dict()
    assert set(a) == set(b)
    lr, steps = 5e-3, trainer.iterations
    for k in a:
        diff = (b[k].double() - a[k].double()).abs()
        if "moving_" in k:
            np.testing.assert_allclose(b[k].cpu().numpy(), a[k].cpu().numpy(), rtol=5e-2, atol=1e-3, err_msg=k)
        else:
            assert float(diff.max()) <= lr * steps and float(diff.mean()) <= 0.05 * lr, (k, float(diff.max()), float(diff.mean()))
