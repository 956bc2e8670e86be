"""The recall-SDPA two-tower forward (fused bags -> SDPA encoder -> tower MLPs -> l2 norm -> in-batch
softmax loss) and TransformerEncoder vs numpy float64 restatements built on the oracle."""
import os

import numpy as np
import pytest
import torch
import yaml

import oracle
from recommendflow_b200.backend.layers.network_layers import TransformerEncoder
from recommendflow_b200.config_parser import Configuration, Features
from recommendflow_b200.models.matching.recall_sdpa import RecallSdpa
from recommendflow_b200.strings import StringColumn

pytestmark = pytest.mark.gpu


def selu(x):
    a, s = 1.6732632423543772, 1.0507009873554805
    return s * np.where(x > 0, x, a * (np.exp(np.minimum(x, 0)) - 1))


def test_recall_sdpa_forward_matches_numpy(golden_dir, monkeypatch):
    import recommendflow_b200.dense_ops as dense_ops
    rng = np.random.default_rng(31)
    conf = Configuration(os.path.join(golden_dir, "configs", "synth_recall_sdpa.yaml"),
                         slot_map_path=os.path.join(golden_dir, "configs", "synth_recall_sdpa.feature.map"))
    # keep the test light: 5 user + 4 ad features of the 228
    keep = set(conf.features.user_feature_names[:5] + conf.features.ad_feature_names[:4])
    for f in conf.features.features:
        if f.is_hashing() and f.name not in keep:
            f.working = False
    B, S, dm = 96, 12, 32
    model = RecallSdpa(conf, tower_units=(48, 24), behaviour_dim=dm, num_heads=2)
    assert len(model.user_cols) == 5 and len(model.ad_cols) == 4
    batch, host = {}, {}
    for name in model.user_cols + model.ad_cols:
        L = int(rng.integers(1, 4))
        rows = [[f"{name[:3]}{rng.integers(0, 50)}" for _ in range(rng.integers(0, L + 1))] for _ in range(B)]
        rows[0] = ["x"] * L
        batch[name] = StringColumn.from_lists(rows).to("cuda")
        layer = model.preprocessor[name]
        ws = [rng.uniform(-0.05, 0.05, size=(layer.num_bins, layer.output_dim)).astype(np.float32) for _ in range(2)]
        layer.set_weights(ws)
        flat = [x for r in rows for x in (r + [""] * (L - len(r)))]
        arena, offs = oracle.encode_strings(flat)
        host[name] = oracle.hashed_bag_forward(arena, offs, B, L, ws, [layer.num_bins] * 2, [2022, 2023], "sum").astype(np.float64)
    x = rng.standard_normal((B, S, dm)).astype(np.float32)
    mask = (np.arange(S)[None, :, None] < rng.integers(1, S + 1, size=(B, 1, 1))).astype(np.float32)
    y = (rng.uniform(size=B) > 0.3).astype(np.float32)
    # explicit weights everywhere
    enc = []
    for dense in (model.seq_encoder.wq, model.seq_encoder.wk, model.seq_encoder.wv):
        w, b = (rng.standard_normal((dm, dm)) * 0.2).astype(np.float32), (rng.standard_normal(dm) * 0.1).astype(np.float32)
        dense.set_weights([w, b])
        enc.append((w, b))
    towers = {}
    for tname, mlp, d_in in (("u", model.user_dense, 5 * 16 + dm), ("a", model.ad_dense, 4 * 16)):
        dims, params = [d_in, 48, 24], []
        dense_layers = [l for l in mlp.layers if hasattr(l, "dense")]
        for li, dl in enumerate(dense_layers):
            w = (rng.standard_normal((dims[li], dims[li + 1])) / np.sqrt(dims[li])).astype(np.float32)
            b = (rng.standard_normal(dims[li + 1]) * 0.1).astype(np.float32)
            dl.dense.set_weights([w, b])
            params.append((w, b))
        towers[tname] = params

    for precision, tol in (("fp32", 2e-4), ("tf32", 2e-2)):
        monkeypatch.setattr(dense_ops, "DEFAULT_PRECISION", precision)
        out = model(batch, y_true=torch.from_numpy(y).cuda(),
                    behaviour=(torch.from_numpy(x).cuda(), torch.from_numpy(mask).cuda()), training=False)
        loss = model(batch, y_true=torch.from_numpy(y).cuda(),
                     behaviour=(torch.from_numpy(x).cuda(), torch.from_numpy(mask).cuda()), training=True)
        # numpy float64 restatement
        proj = [(x.astype(np.float64) @ w + b).astype(np.float32) for w, b in enc]
        heads = [p.reshape(B, S, 2, dm // 2).transpose(0, 2, 1, 3) for p in proj]
        att = oracle.sdpa(*heads, np.broadcast_to(mask[:, None], (B, 2, S, 1))).transpose(0, 2, 1, 3).reshape(B, S, dm)
        u = np.concatenate([host[n] for n in model.user_cols] + [att.mean(axis=1).astype(np.float64)], axis=1)
        a = np.concatenate([host[n] for n in model.ad_cols], axis=1)

        def tower(v, params):
            for w, b in params:
                v = selu((v / np.sqrt(1.0 + 1e-6)) @ w.astype(np.float64) + b)     # BatchNorm with fresh moving stats
            return v / np.maximum(np.linalg.norm(v, axis=1, keepdims=True), 1e-12)
        u, a = tower(u, towers["u"]), tower(a, towers["a"])
        np.testing.assert_allclose(out["user"].cpu().numpy(), u, rtol=1e-3, atol=tol)
        # no SDPA on the ad side; in tf32 mode the tower GEMMs go through cuBLAS' TF32 path too
        np.testing.assert_allclose(out["ad"].cpu().numpy(), a, rtol=1e-3, atol=2e-5 if precision == "fp32" else tol)
        want, _, _ = oracle.inbatch_softmax_ce(y, u.astype(np.float32), a.astype(np.float32), 20.0)
        assert abs(float(loss) - want) <= max(tol * 20, 5e-3), (precision, float(loss), want)


def test_transformer_encoder_keras_semantics(monkeypatch):
    import recommendflow_b200.dense_ops as dense_ops
    monkeypatch.setattr(dense_ops, "DEFAULT_PRECISION", "fp32")
    rng = np.random.default_rng(32)
    B, S, d, H, hid = 5, 9, 16, 2, 24          # Keras MHA gets (num_heads=d, key_dim=H): 16 heads of size 2
    x = rng.standard_normal((B, S, d)).astype(np.float32)
    mask = (np.arange(S)[None, :, None] < rng.integers(1, S + 1, size=(B, 1, 1))).astype(np.float32)
    layer = TransformerEncoder(d, num_heads=H, ffn_hidden_unit=hid)
    assert (layer.mha.num_heads, layer.mha.key_dim) == (d, H)
    N = d
    W = {n: (rng.standard_normal(s) * 0.3).astype(np.float32) for n, s in
         [("wq", (d, N, H)), ("bq", (N, H)), ("wk", (d, N, H)), ("bk", (N, H)), ("wv", (d, N, H)), ("bv", (N, H)),
          ("wo", (N, H, d)), ("bo", (d,))]}
    layer.mha.set_weights([W[n] for n in ("wq", "bq", "wk", "bk", "wv", "bv", "wo", "bo")])
    w1, b1 = (rng.standard_normal((d, hid)) * 0.3).astype(np.float32), (rng.standard_normal(hid) * 0.1).astype(np.float32)
    w2, b2 = (rng.standard_normal((hid, d)) * 0.3).astype(np.float32), (rng.standard_normal(d) * 0.1).astype(np.float32)
    layer.ffn.conv1.set_weights([w1, b1])
    layer.ffn.conv2.set_weights([w2, b2])
    got = layer([torch.from_numpy(x).cuda(), torch.from_numpy(mask).cuda()]).cpu().numpy()

    xd = x.astype(np.float64)
    q = np.einsum("abc,cde->abde", xd, W["wq"]) + W["bq"]
    k = np.einsum("abc,cde->abde", xd, W["wk"]) + W["bk"]
    v = np.einsum("abc,cde->abde", xd, W["wv"]) + W["bv"]
    s = np.einsum("aecd,abcd->acbe", k, q / np.sqrt(H))            # [B, N, T, S]
    # Keras' additive mask, broadcast over keys -- in float32, as Keras computes it: x + (-1e9) rounds to
    # -1e9 exactly (ulp 64), so a masked QUERY row attends uniformly
    s = (s.astype(np.float32) + ((1.0 - mask[:, None, :, :]) * -1e9).astype(np.float32)).astype(np.float64)
    p = np.exp(s - s.max(-1, keepdims=True))
    p /= p.sum(-1, keepdims=True)
    ctx = np.einsum("acbe,aecd->abcd", p, v)
    att = np.einsum("abcd,cde->abe", ctx, W["wo"]) + W["bo"]

    def ln(z):
        mu, var = z.mean(-1, keepdims=True), z.var(-1, keepdims=True)
        return (z - mu) / np.sqrt(var + 1e-6)
    out1 = ln(xd + att)
    want = ln(out1 + (np.maximum(out1 @ w1 + b1, 0) @ w2 + b2))
    np.testing.assert_allclose(got, want, rtol=1e-3, atol=1e-4)
    # tensor-core mode: the four projections and the two FFN layers run on rf_dense_forward_tc (TF32 operands); the
    # values pass through two LayerNorms, so the TF32 operand rounding (2^-10 relative) shows up at the 1e-2 level
    from recommendflow_b200 import _native as nat
    monkeypatch.setattr(dense_ops, "DEFAULT_PRECISION", "tf32")
    before = nat.launch_count()
    got_tc = layer([torch.from_numpy(x).cuda(), torch.from_numpy(mask).cuda()]).cpu().numpy()
    assert nat.launch_count() - before == 7, "q, k, v, o projections + SDPA + two FFN layers"
    np.testing.assert_allclose(got_tc, want, rtol=3e-2, atol=3e-2)
