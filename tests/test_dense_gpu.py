"""SDPA, MultiHeadAttention and the in-batch two-tower losses (CUDA, through the C-ABI) vs the oracle.

Floating point: the kernels accumulate in fp32, the oracle in float64.  Tolerances (stated per test)
cover fp32 rounding of length-D dot products: |err| <~ D * 2^-24 * |q||k| on a logit, times the
temperature (20) inside the exponent for the losses."""
import numpy as np
import pytest
import torch

import oracle
from recommendflow_b200.backend.layers.attention_layers import MultiHeadAttention, SelfAttention
from recommendflow_b200.backend.layers.layer_utils import scaled_dot_product_attention, split_heads
from recommendflow_b200.backend.lossess import match_losses, match_zipped_losses
from recommendflow_b200 import _native as nat
from recommendflow_b200.dense_ops import inbatch_rowstats
from recommendflow_b200.utils.str_parser import str2loss

pytestmark = pytest.mark.gpu


# tf32 = tcgen05 path (S <= 64, dh in 32/64/96; other shapes fall back to the fp32 kernel): the tensor core
# truncates the fp32 operands to TF32 (2^-10 relative), logits of N(0,1) inputs move by up to ~1e-2.
SDPA_TOL = {"fp32": dict(rtol=2e-5, atol=2e-6), "tf32": dict(rtol=2e-2, atol=2e-2)}


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
@pytest.mark.parametrize("shape", [(7, 50, 64), (3, 4, 50, 16), (2, 1, 1), (5, 2, 33, 8), (4, 200, 32), (16, 50, 128),
                                   (9, 50, 32), (2, 3, 64, 96), (301, 17, 64), (1, 1, 32)])
def test_sdpa_matches_oracle(shape, precision, monkeypatch):
    import recommendflow_b200.dense_ops as dense_ops
    monkeypatch.setattr(dense_ops, "DEFAULT_PRECISION", precision)
    tol = SDPA_TOL[precision]
    rng = np.random.default_rng(sum(shape))
    q, k, v = (rng.standard_normal(shape).astype(np.float32) for _ in range(3))
    mask = (rng.uniform(size=shape[:-1] + (1,)) > 0.3).astype(np.float32)
    mask[0, ..., :, :] = 0                                   # a fully masked sequence -> uniform attention
    want = oracle.sdpa(q, k, v, mask)
    got = scaled_dot_product_attention(*(torch.from_numpy(x).cuda() for x in (q, k, v, mask))).cpu().numpy()
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, **tol)
    no_mask = scaled_dot_product_attention(*(torch.from_numpy(x).cuda() for x in (q, k, v)), None).cpu().numpy()
    np.testing.assert_allclose(no_mask, oracle.sdpa(q, k, v, None), **tol)
    # masked query rows attend uniformly: output = mean over keys of v
    np.testing.assert_allclose(got[0], np.broadcast_to(v[0].mean(axis=-2, keepdims=True), v[0].shape),
                               rtol=max(tol["rtol"], 1e-5), atol=max(tol["atol"] / 10, 1e-6))


def test_multi_head_attention_layer_matches_reference_semantics(monkeypatch):
    import recommendflow_b200.dense_ops as dense_ops
    monkeypatch.setattr(dense_ops, "DEFAULT_PRECISION", "fp32")
    rng = np.random.default_rng(9)
    B, S, d_model, H = 6, 50, 64, 4
    x = rng.standard_normal((B, S, d_model)).astype(np.float32)
    mask = (np.arange(S)[None, :, None] < rng.integers(1, S + 1, size=(B, 1, 1))).astype(np.float32)
    layer = MultiHeadAttention(d_model, H)
    ws = []
    for dense in (layer.wq, layer.wk, layer.wv):
        w = (rng.standard_normal((d_model, d_model)) * 0.1).astype(np.float32)
        b = (rng.standard_normal(d_model) * 0.1).astype(np.float32)
        dense.set_weights([w, b])
        ws.append((w, b))
    xt, mt = torch.from_numpy(x).cuda(), torch.from_numpy(mask).cuda()
    got = layer(xt, xt, xt, mt).cpu().numpy()
    proj = [(x.astype(np.float64) @ w + b).astype(np.float32) for w, b in ws]
    heads = [p.reshape(B, S, H, d_model // H).transpose(0, 2, 1, 3) for p in proj]
    att = oracle.sdpa(*heads, np.broadcast_to(mask[:, None], (B, H, S, 1)))
    want = att.transpose(0, 2, 1, 3).reshape(B, S, d_model)
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5)          # + fp32 cuBLAS projections
    assert split_heads(xt, S, H, d_model // H).shape == (B, H, S, d_model // H)
    sa = SelfAttention(add_pos=True)
    assert sa([xt, xt, xt, mt]).shape == (B, d_model)


def _pairs(rng, B, D):
    q = rng.standard_normal((B, D)).astype(np.float32)
    d = (0.7 * q + 0.7 * rng.standard_normal((B, D))).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    y = (rng.uniform(size=B) > 0.2).astype(np.float32)
    return q, d, y


# fp32: |S| <= 1, fp32 dot of <= 256 terms: ~1e-6 on a logit, x 20 in the exponent -> 1e-4.
# tf32 (tcgen05): operands rounded to nearest TF32 (2^-11 relative each, unbiased): |dS| <= 2 * 2^-11 *
# sum_k |q_k d_k| <= 1e-3 for unit vectors (worst case, all errors aligned), x 20 in the exponent ->
# 2e-2 bound on a row's lse; observed ~4e-3 at D = 16 and ~1e-3 at D = 256.
# bf16 operands (kind::f16): round-to-nearest to 8 mantissa bits, <= 2^-9 relative per operand: |dS_ij| <= 2^-8 for unit
# vectors in the worst case (all errors aligned), x 20 in the exponent -> 8e-2 bound on a row's lse; observed ~5e-3.
TOL = {"fp32": 1e-4, "tf32": 2e-2, "bf16": 8e-2}


@pytest.mark.parametrize("precision", ["fp32", "tf32", "bf16"])
@pytest.mark.parametrize("B,D", [(1000, 256), (64, 16), (777, 100), (129, 8), (4096, 256), (300, 36), (8192, 64), (255, 128), (257, 24)])
def test_scaled_inbatch_softmax_loss_matches_oracle(B, D, precision, monkeypatch):
    import recommendflow_b200.dense_ops as dense_ops
    monkeypatch.setattr(dense_ops, "DEFAULT_PRECISION", "tf32" if precision == "bf16" else precision)
    monkeypatch.setattr(dense_ops, "DEFAULT_LOSS_PRECISION", precision)
    rng = np.random.default_rng(B + D)
    q, d, y = _pairs(rng, B, D)
    want, lse, diag = oracle.inbatch_softmax_ce(y, q, d, 20.0)
    qt, dt, yt = (torch.from_numpy(a).cuda() for a in (q, d, y))
    got = match_losses.batch_neg_sample_scaled_multi_class_ce_loss(yt, qt, dt)
    r = inbatch_rowstats(qt, dt, scale=20.0, want=("lse", "diag"))
    np.testing.assert_allclose(r["diag"].cpu().numpy(), diag, atol=2e-6)          # the diagonal is exact fp32 in both modes
    np.testing.assert_allclose(r["lse"].cpu().numpy(), lse, atol=TOL[precision])
    assert abs(float(got) - want) <= TOL[precision]
    if precision != "fp32":
        return
    z = match_zipped_losses.batch_neg_sample_scaled_multi_class_ce_loss(
        yt[:, None], match_zipped_losses.zip_embedding(qt * 3.0, dt * 0.5))       # wrapper re-normalises
    assert abs(float(z) - want) <= 1e-4
    sym = match_losses.batch_neg_sample_symmetrical_scaled_multi_class_ce_loss(yt, qt, dt, scale=3)
    want_sym, _, _ = oracle.inbatch_softmax_ce(y, q, d, 9.0)             # reference applies `scale` twice
    assert abs(float(sym) - want_sym) <= 1e-4


@pytest.mark.parametrize("precision", ["fp32", "tf32", "bf16"])
def test_margin_rank_losses_match_numpy(precision, monkeypatch):
    import recommendflow_b200.dense_ops as dense_ops
    monkeypatch.setattr(dense_ops, "DEFAULT_PRECISION", "tf32" if precision == "bf16" else precision)
    monkeypatch.setattr(dense_ops, "DEFAULT_LOSS_PRECISION", precision)
    rng = np.random.default_rng(21)
    B, D = 513, 64
    q, d, y = _pairs(rng, B, D)
    S = q.astype(np.float64) @ d.astype(np.float64).T
    qt, dt, yt = (torch.from_numpy(a).cuda() for a in (q, d, y))
    want = (np.clip(-(np.diag(S)[:, None] - S) + 0.1, 0, 1e14) * y[None, :]).sum()      # y broadcasts over columns
    got = float(match_losses.batch_neg_sample_margin_rank_loss(yt, qt, dt, margin=0.1))
    rel = {"fp32": 1e-3, "tf32": 5e-3, "bf16": 2e-2}[precision]
    assert abs(got - want) <= rel * max(1.0, abs(want))
    neg = (S - np.diag(np.diag(S))).max(axis=-1)
    want_h = (np.clip(-(np.diag(S) - neg) + 0.1, 0, 1e14) * y).sum()
    got_h = float(match_losses.batch_hard_neg_sample_margin_rank_loss(yt, qt, dt, margin=0.1))
    assert abs(got_h - want_h) <= {"fp32": 1e-4, "tf32": 5e-3, "bf16": 2e-2}[precision] * max(1.0, abs(want_h))
    mse = float(match_losses.mean_squared_error(yt, qt, dt))
    assert abs(mse - np.mean((y - np.diag(S)) ** 2)) <= 1e-5


def test_loss_lookup_by_dotted_name_and_initials():
    f = str2loss("backend.losses.match_losses.bnssmccl")
    assert f is match_losses.batch_neg_sample_scaled_multi_class_ce_loss
    assert str2loss("backend.lossess.match_losses.cosent_loss") is match_losses.cosent_loss


def test_full_size_logits_properties():
    # C3 size: B = 8192, Dt = 256.  Property: with doc == query (unit rows) S_ii = 1 is each row's max, so
    # lse_i >= 20 and the loss is >= 0; swapping two docs changes exactly the losses of those rows.
    torch.manual_seed(0)
    B, D = 8192, 256
    q = torch.nn.functional.normalize(torch.randn(B, D, device="cuda"), dim=1)
    y = torch.ones(B, device="cuda")
    rows = torch.randint(0, B, (256,), device="cuda")
    ref = torch.logsumexp(20.0 * (q[rows].double() @ q.double().T), dim=1)
    for precision, tol in TOL.items():
        r = inbatch_rowstats(q, q, y_true=y, scale=20.0, want=("lse", "diag"), precision=precision)
        assert torch.allclose(r["diag"], torch.ones(B, device="cuda"), atol=1e-5)
        assert float(r["lse"].min()) >= 20.0 - tol and float(r["loss"]) >= -tol
        assert torch.allclose(r["lse"][rows].double(), ref, atol=tol), precision


# ---- backward of the dense contractions: CUDA kernels vs torch.autograd on a float64 restatement -----------
def _ref_sdpa64(q, k, v, mask):
    s = q @ k.transpose(-1, -2) / np.sqrt(q.shape[-1])
    if mask is not None:
        s = torch.where(mask == 0, torch.full_like(s, -4294967295.0), s)      # [..., S, 1] broadcasts over keys
    return torch.softmax(s, dim=-1) @ v


@pytest.mark.parametrize("shape", [(6, 50, 32), (3, 2, 50, 64), (2, 64, 128), (5, 1, 8), (4, 3, 17, 20)])
def test_sdpa_backward_matches_autograd(shape):
    from recommendflow_b200.dense_ops import sdpa_autograd
    rng = np.random.default_rng(sum(shape) + 1)
    q, k, v, g = (torch.from_numpy(rng.standard_normal(shape).astype(np.float32)).cuda() for _ in range(4))
    mask = torch.from_numpy((rng.uniform(size=shape[:-1] + (1,)) > 0.3).astype(np.float32)).cuda()
    mask[0] = 0                                                               # a fully masked sequence
    for m in (mask, None):
        q64, k64, v64 = (t.double().requires_grad_(True) for t in (q, k, v))
        _ref_sdpa64(q64, k64, v64, None if m is None else m.double()).backward(g.double())
        qg, kg, vg = (t.clone().requires_grad_(True) for t in (q, k, v))
        out = sdpa_autograd(qg, kg, vg, m, "fp32")
        out.backward(g)
        for name, got, want in (("dq", qg.grad, q64.grad), ("dk", kg.grad, k64.grad), ("dv", vg.grad, v64.grad)):
            # fp32 accumulation of <= 128-term dot products of N(0,1) values, chained three deep
            np.testing.assert_allclose(got.cpu().numpy(), want.float().cpu().numpy(), rtol=2e-4, atol=2e-5, err_msg=name)
        if m is not None:                                                     # masked query rows pass nothing to q
            assert float(qg.grad[0].abs().max()) == 0.0
    with pytest.raises(NotImplementedError):
        big = torch.zeros(1, 65, 8, device="cuda", requires_grad=True)
        sdpa_autograd(big, big, big, None, "fp32").sum().backward()


@pytest.mark.parametrize("shape", [(37, 50, 64), (3, 2, 50, 32), (5, 64, 128), (9, 17, 96), (4, 1, 64)])
def test_sdpa_backward_tensor_cores_match_float64(shape):
    """rf_sdpa_backward_tc: the five products of the attention backward on warp-level TF32 tensor-core MMAs (transposed operands
    read out of shared memory), fp32 softmax / delta; vs torch autograd on a float64 restatement and vs the exact-fp32 kernel."""
    from recommendflow_b200.dense_ops import sdpa_backward
    rng = np.random.default_rng(sum(shape))
    q, k, v, g = (torch.from_numpy(rng.standard_normal(shape).astype(np.float32)).cuda() for _ in range(4))
    mask = torch.from_numpy((rng.uniform(size=shape[:-1] + (1,)) > 0.3).astype(np.float32)).cuda()
    mask[0] = 0
    for m in (mask, None):
        q64, k64, v64 = (t.double().requires_grad_(True) for t in (q, k, v))
        _ref_sdpa64(q64, k64, v64, None if m is None else m.double()).backward(g.double())
        before = nat.launch_count()
        got = sdpa_backward(q, k, v, m, g, precision="tf32")
        assert nat.launch_count() == before + 1
        exact = sdpa_backward(q, k, v, m, g, precision="fp32")
        for name, a, e, want in zip(("dq", "dk", "dv"), got, exact, (q64.grad, k64.grad, v64.grad)):
            want = want.float().cpu().numpy()
            # TF32 operands (2^-11 relative each) through three chained products of up to 128 terms
            tol = 6e-3 * max(1.0, float(np.abs(want).max()))
            np.testing.assert_allclose(a.cpu().numpy(), want, rtol=2e-2, atol=tol, err_msg=name)
            np.testing.assert_allclose(a.cpu().numpy(), e.cpu().numpy(), rtol=2e-2, atol=tol, err_msg=name + " vs fp32 kernel")
        if m is not None:
            assert float(got[0][0].abs().max()) == 0.0                         # a fully masked sequence passes nothing to q


@pytest.mark.parametrize("B,D,diag", [(512, 64, True), (4100, 64, True), (2048, 256, False), (6144, 128, True)])
def test_inbatch_softmax_ce_backward_tensor_core_slabs(B, D, diag):
    """rf_inbatch_softmax_ce_backward_tc (three tcgen05 GEMMs per slab of 2048 query rows, ragged last slab) against the
    CUDA-core fp32 kernel on the same lse."""
    from recommendflow_b200.dense_ops import inbatch_rowstats, inbatch_softmax_ce_backward
    g = torch.Generator(device="cuda").manual_seed(B + D)
    q = torch.nn.functional.normalize(torch.randn(B, D, device="cuda", generator=g), dim=1)
    d = torch.nn.functional.normalize(0.6 * q + 0.8 * torch.randn(B, D, device="cuda", generator=g), dim=1)
    y = (torch.rand(B, device="cuda", generator=g) > 0.3).float()
    lse = inbatch_rowstats(q, d, scale=20.0, want=("lse",), precision="fp32")["lse"]
    rq, rd = inbatch_softmax_ce_backward(q, d, y, lse, 20.0, 2.0, positives_on_diagonal=diag, precision="fp32")
    before = nat.launch_count()
    tq, td = inbatch_softmax_ce_backward(q, d, y, lse, 20.0, 2.0, positives_on_diagonal=diag, precision="tf32")
    assert nat.launch_count() > before
    # TF32 operands: relative error ~2^-10 per product on gradients whose rows have norm <= 2 * 20 * 2 / B
    scale = float(rq.abs().max())
    np.testing.assert_allclose(tq.cpu().numpy(), rq.cpu().numpy(), rtol=2e-2, atol=4e-3 * scale)
    np.testing.assert_allclose(td.cpu().numpy(), rd.cpu().numpy(), rtol=2e-2, atol=4e-3 * float(rd.abs().max()))


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
@pytest.mark.parametrize("B,D", [(300, 64), (1000, 256), (77, 20), (2048, 128), (130, 512)])
def test_inbatch_softmax_ce_backward_matches_autograd(B, D, precision):
    from recommendflow_b200.dense_ops import inbatch_softmax_ce_autograd
    rng = np.random.default_rng(B + D)
    q = torch.nn.functional.normalize(torch.from_numpy(rng.standard_normal((B, D)).astype(np.float32)), dim=1).cuda()
    d = torch.nn.functional.normalize(torch.from_numpy(rng.standard_normal((B, D)).astype(np.float32)), dim=1).cuda()
    y = torch.from_numpy((rng.uniform(size=B) > 0.3).astype(np.float32)).cuda()
    q64, d64 = q.double().requires_grad_(True), d.double().requires_grad_(True)
    s = 20.0 * (q64 @ d64.T)
    ref = (-(torch.diagonal(s) - torch.logsumexp(s, dim=1)) * y.double()).mean()
    (3.0 * ref).backward()
    qg, dg = q.clone().requires_grad_(True), d.clone().requires_grad_(True)
    loss = inbatch_softmax_ce_autograd(y, qg, dg, 20.0, precision)
    (3.0 * loss).backward()
    # the gradient is a softmax-weighted sum of unit vectors times 20 / B; in tf32 mode only the forward's lse
    # carries TF32 error (|d lse| <~ 20 * 2^-10), the backward recomputes the logits in fp32
    tol = dict(rtol=1e-3, atol=2e-6) if precision == "fp32" else dict(rtol=3e-2, atol=3e-2 * 60.0 / B)
    assert abs(float(loss) - float(ref)) <= (1e-4 if precision == "fp32" else 3e-2)
    np.testing.assert_allclose(qg.grad.cpu().numpy(), q64.grad.float().cpu().numpy(), **tol)
    np.testing.assert_allclose(dg.grad.cpu().numpy(), d64.grad.float().cpu().numpy(), **tol)
    # only one side requested
    qh = q.clone().requires_grad_(True)
    inbatch_softmax_ce_autograd(y, qh, d, 20.0, "fp32").backward()
    np.testing.assert_allclose(qh.grad.cpu().numpy(), (q64.grad / 3.0).float().cpu().numpy(), rtol=1e-3, atol=2e-6)


# ---- every two-tower loss is connected to the autograd graph (only the scaled CE was) ---------------
def _ref_losses64():
    """float64 torch restatements of /root/reference/backend/lossess/match_losses.py (B x B materialised)."""
    def S(q, d):
        return q @ d.t()

    def pairwise(s, y, positive_only):
        diff = s[:, None] - s[None, :]
        keep = y[:, None] < y[None, :]
        if positive_only:
            keep = keep & (diff > 0)
        flat = torch.where(keep, diff, torch.full_like(diff, -1e12)).reshape(-1)
        return torch.logsumexp(torch.cat([torch.zeros(1, dtype=s.dtype, device=s.device), flat]), dim=0)

    def ce(y, q, d):
        s = S(q, d)
        p = torch.clamp(torch.diagonal(s) / s.sum(dim=1), 1e-7, 1 - 1e-7)
        return torch.mean(-y * torch.log(p) * y)

    def sym_ce(y, q, d):
        s = S(q, d)
        p1 = torch.clamp(torch.diagonal(s) / s.sum(dim=1), 1e-7, 1 - 1e-7)
        p2 = torch.clamp(torch.diagonal(s) / s.sum(dim=0), 1e-7, 1 - 1e-7)
        return torch.mean(0.5 * (-y * torch.log(p1) - y * torch.log(p2)) * y)

    def scaled(y, q, d, scale=20.0):
        s = S(q, d) * scale
        return torch.mean(-(torch.diagonal(s) - torch.logsumexp(s, dim=1)) * y)

    def margin_rank(y, q, d, margin=0.1):
        s = S(q, d)
        return (torch.clamp(s - torch.diagonal(s)[:, None] + margin, 0, 1e14) * y[None, :]).sum()

    def hard_rank(y, q, d, margin=0.1):
        s = S(q, d)
        neg = s - torch.diag(torch.diagonal(s))
        return (torch.clamp(neg.max(dim=1).values - torch.diagonal(s) + margin, 0, 1e14) * y).sum()

    diag = lambda q, d: (q * d).sum(dim=1)
    return {
        "mean_squared_error": lambda y, q, d: torch.mean((y - diag(q, d)) ** 2),
        "binary_cross_entropy": lambda y, q, d: (-(y * torch.log(torch.clamp(diag(q, d), 1e-7, 1 - 1e-7)) + (1 - y) * torch.log(
            1 - torch.clamp(diag(q, d), 1e-7, 1 - 1e-7)))).sum(),
        "cosent_loss": lambda y, q, d: pairwise(diag(q, d) * 20, y, False),
        "cosent_loss_v2": lambda y, q, d: pairwise(diag(q, d) * 20, y, True),
        "batch_neg_sample_ce_loss": ce,
        "batch_neg_sample_symmetrical_ce_loss": sym_ce,
        "batch_neg_sample_scaled_multi_class_ce_loss": scaled,
        "batch_neg_sample_symmetrical_scaled_multi_class_ce_loss": lambda y, q, d: scaled(y, q, d, 400.0),
        "batch_neg_sample_margin_rank_loss": margin_rank,
        "batch_hard_neg_sample_margin_rank_loss": hard_rank,
    }


@pytest.mark.parametrize("name", sorted(_ref_losses64()))
def test_every_loss_value_and_gradient_match_float64_autograd(name):
    """loss.backward() reaches query AND doc through every term of the loss (rtol 2e-3 on the gradients: fp32 / TF32
    contractions against float64; the symmetric scaled loss runs at temperature 400, where TF32 logits move the
    softmax visibly, so it is checked on the exact-fp32 kernel)."""
    import recommendflow_b200.dense_ops as dense_ops
    rng = np.random.default_rng(len(name))
    B, D = 300, 64
    q = rng.standard_normal((B, D)); d = rng.standard_normal((B, D))
    if name in ("batch_neg_sample_ce_loss", "batch_neg_sample_symmetrical_ce_loss", "binary_cross_entropy"):
        q, d = np.abs(q), np.abs(d)                      # keep the "probabilities" positive like trained towers would
    q /= np.linalg.norm(q, axis=1, keepdims=True); d /= np.linalg.norm(d, axis=1, keepdims=True)
    y = (rng.uniform(size=B) > 0.4).astype(np.float64) if "cosent" not in name else rng.integers(0, 4, size=B).astype(np.float64)
    q64 = torch.tensor(q, dtype=torch.float64, device="cuda", requires_grad=True)
    d64 = torch.tensor(d, dtype=torch.float64, device="cuda", requires_grad=True)
    want = _ref_losses64()[name](torch.tensor(y, device="cuda"), q64, d64)
    want.backward()
    q32 = torch.tensor(q, dtype=torch.float32, device="cuda", requires_grad=True)
    d32 = torch.tensor(d, dtype=torch.float32, device="cuda", requires_grad=True)
    old = dense_ops.DEFAULT_PRECISION
    dense_ops.DEFAULT_PRECISION = "fp32"
    try:
        got = getattr(match_losses, name)(torch.tensor(y, dtype=torch.float32, device="cuda"), q32, d32)
        if got.dim():                                     # binary_cross_entropy returns the per-row vector (:35-39)
            got = got.sum()
        got.backward()
    finally:
        dense_ops.DEFAULT_PRECISION = old
    assert q32.grad is not None and d32.grad is not None, "loss is detached from query / doc"
    np.testing.assert_allclose(got.item(), want.item(), rtol=2e-4, atol=1e-5)
    scale = max(q64.grad.abs().max().item(), d64.grad.abs().max().item(), 1e-12)
    np.testing.assert_allclose(q32.grad.double().cpu().numpy() / scale, q64.grad.cpu().numpy() / scale, rtol=0, atol=2e-3)
    np.testing.assert_allclose(d32.grad.double().cpu().numpy() / scale, d64.grad.cpu().numpy() / scale, rtol=0, atol=2e-3)


def test_aux_label_cosent_losses_match_reference_formula():
    """aux_label_cosent_loss / pos_aux_label_cosent_loss (match_losses.py:72-116): cosent_v2 over the positive
    (and negative) subsets of the batch, on the auxiliary label."""
    rng = np.random.default_rng(77)
    B, D = 64, 16
    q = rng.standard_normal((B, D)); d = rng.standard_normal((B, D))
    q /= np.linalg.norm(q, axis=1, keepdims=True); d /= np.linalg.norm(d, axis=1, keepdims=True)
    y = (rng.uniform(size=B) > 0.5).astype(np.float32)
    aux = rng.uniform(size=B).astype(np.float32)

    def v2(idx):
        s = (q[idx] * d[idx]).sum(axis=1) * 20
        diff = s[:, None] - s[None, :]
        keep = (aux[idx][:, None] < aux[idx][None, :]) & (diff > 0)
        return np.log(1.0 + np.exp(diff[keep]).sum())

    pos, neg = np.nonzero(y == 1)[0], np.nonzero(y == 0)[0]
    args = [torch.tensor(x, dtype=torch.float32, device="cuda") for x in (y, aux, q, d)]
    np.testing.assert_allclose(match_losses.pos_aux_label_cosent_loss(*args).item(), v2(pos), rtol=1e-4)
    np.testing.assert_allclose(match_losses.aux_label_cosent_loss(*args, alpha=0.3).item(), 0.7 * v2(pos) + 0.3 * v2(neg), rtol=1e-4)
    assert match_losses.pos_aux_label_cosent_loss(torch.zeros(B, device="cuda"), *args[1:]).item() == 0.0


# ---- Keras Dense / tower MLP on the tcgen05 GEMM (rf_dense_forward_tc) ----------------------------------------------
def _act64(name, z):
    if name in (None, "linear"):
        return z
    if name == "relu":
        return np.maximum(z, 0)
    if name == "selu":
        return 1.0507009873554805 * np.where(z > 0, z, 1.6732632423543772 * np.expm1(z))
    if name == "tanh":
        return np.tanh(z)
    if name == "sigmoid":
        return 1 / (1 + np.exp(-z))
    if name == "gelu":
        from math import erf
        return 0.5 * z * (1 + np.vectorize(erf)(z / np.sqrt(2)))
    raise AssertionError(name)


# TF32 operands (the tensor core drops the low 13 mantissa bits of each fp32 operand, 2^-10 relative), fp32 accumulate:
# |err| <~ 2^-10 * sum_k |x_k w_k|; for the sizes below that is a few 1e-3
DENSE_TOL = dict(rtol=6e-3, atol=6e-3)


@pytest.mark.parametrize("M,K,N,act", [(8192, 1664, 1024, "selu"), (300, 64, 64, None), (129, 512, 256, "relu"), (1000, 1024, 512, "tanh"),
                                       (77, 36, 20, "sigmoid"), (4096, 256, 128, "gelu"), (1, 32, 4, None), (513, 100, 260, "selu")])
def test_dense_forward_tc_matches_float64(M, K, N, act):
    from recommendflow_b200.dense_ops import dense_forward
    rng = np.random.default_rng(M + K + N)
    x = rng.standard_normal((M, K)).astype(np.float32)
    w = rng.uniform(-0.05, 0.05, size=(K, N)).astype(np.float32)          # Keras kernel [in, units]
    b = rng.uniform(-0.5, 0.5, size=N).astype(np.float32)
    want = _act64(act, x.astype(np.float64) @ w.astype(np.float64) + b)
    before = __import__("recommendflow_b200")._native.launch_count()
    got = dense_forward(torch.from_numpy(x).cuda(), torch.from_numpy(np.ascontiguousarray(w.T)).cuda(), torch.from_numpy(b).cuda(), act)
    assert __import__("recommendflow_b200")._native.launch_count() == before + 1
    np.testing.assert_allclose(got.cpu().numpy(), want, **DENSE_TOL)
    # no bias; rows l2-normalised in the epilogue (units <= 256)
    if N <= 256:
        z = _act64(act, x.astype(np.float64) @ w.astype(np.float64))
        want_n = z / np.maximum(np.linalg.norm(z, axis=1, keepdims=True), 1e-12)
        got_n = dense_forward(torch.from_numpy(x).cuda(), torch.from_numpy(np.ascontiguousarray(w.T)).cuda(), None, act, l2_normalize=True)
        np.testing.assert_allclose(got_n.cpu().numpy(), want_n, rtol=6e-3, atol=2e-3)


def test_dense_forward_tc_strided_input_and_output_slot():
    """x may be a column window of a wider buffer (the fused bag output), out a column window of another."""
    from recommendflow_b200.dense_ops import dense_forward
    rng = np.random.default_rng(5)
    M, K, N = 700, 96, 64
    wide = rng.standard_normal((M, 256)).astype(np.float32)
    w = rng.uniform(-0.1, 0.1, size=(K, N)).astype(np.float32)
    xw = torch.from_numpy(wide).cuda()
    outw = torch.full((M, 200), 7.0, device="cuda")
    dense_forward(xw[:, 32:32 + K], torch.from_numpy(np.ascontiguousarray(w.T)).cuda(), None, "relu", out=outw[:, 100:100 + N])
    want = np.maximum(wide[:, 32:32 + K].astype(np.float64) @ w, 0)
    np.testing.assert_allclose(outw[:, 100:100 + N].cpu().numpy(), want, **DENSE_TOL)
    assert torch.all(outw[:, :100] == 7.0) and torch.all(outw[:, 100 + N:] == 7.0)      # nothing outside the slot is touched


@pytest.mark.parametrize("rows,in_dim,units,act", [(512, 96, 256, "selu"), (300, 1000, 516, None), (2048, 1888, 1024, "relu"),
                                                   (257, 64, 132, "tanh"), (8192, 512, 256, None)])
def test_dense_forward_cta_pairs_match_float64(rows, in_dim, units, act, monkeypatch):
    """The cta_group::2 variant of the Dense kernel (a cluster of two CTAs computes a 256-row tile; each stages its 128 rows of x
    and half of the weight tile, the leader issues tcgen05.mma.cta_group::2 and a multicast commit frees both rings), forced on
    (RF_DENSE_PAIR=2) for shapes with ragged last row pairs / column tiles, and compared with the single-CTA kernel."""
    from recommendflow_b200.dense_ops import dense_forward
    g = torch.Generator(device="cuda").manual_seed(rows + units)
    x = torch.randn(rows, in_dim, device="cuda", generator=g)
    wt = torch.randn(units, in_dim, device="cuda", generator=g) / in_dim ** 0.5
    b = torch.randn(units, device="cuda", generator=g) * 0.1
    monkeypatch.setenv("RF_DENSE_PAIR", "0")
    single = dense_forward(x, wt, b, act)
    monkeypatch.setenv("RF_DENSE_PAIR", "2")
    pair = dense_forward(x, wt, b, act)
    want = _act64(act, x.double().cpu().numpy() @ wt.double().cpu().numpy().T + b.double().cpu().numpy())
    np.testing.assert_allclose(pair.cpu().numpy(), want, rtol=2e-2, atol=4e-3)
    # same operands, same TF32 products, same accumulation order per output element: the two kernels agree bit for bit
    assert torch.equal(pair, single)
    assert torch.equal(pair, dense_forward(x, wt, b, act))


@pytest.mark.parametrize("rows,in_dim,units", [(512, 8192, 256), (64, 40960, 192), (1888, 8192, 1024), (100, 1000, 68)])
def test_dense_forward_split_k_matches_float64(rows, in_dim, units):
    """A plain product with few output tiles and a long contraction (dW = X^T dZ) splits K over CTAs and sums the partial
    products in order (rf_dense_forward_tc_ex); shapes that do not qualify take the single-pass kernel through the same entry."""
    from recommendflow_b200.dense_ops import dense_forward
    g = torch.Generator(device="cuda").manual_seed(rows + units)
    x = torch.randn(rows, in_dim, device="cuda", generator=g)
    wt = torch.randn(units, in_dim, device="cuda", generator=g)
    split = int(nat.lib().rf_dense_tc_workspace_bytes(rows, in_dim, units)) > 0
    if (rows, in_dim, units) in [(512, 8192, 256), (64, 40960, 192)]:
        assert split, "few output tiles + a long contraction must split K"
    before = nat.launch_count()
    got = dense_forward(x, wt)
    assert nat.launch_count() - before == (2 if split else 1)
    want = x.double() @ wt.double().t()
    # TF32 operands (2^-11 relative each), fp32 accumulation over in_dim terms of size ~1
    np.testing.assert_allclose(got.cpu().numpy(), want.cpu().numpy(), rtol=2e-2, atol=2e-3 * in_dim ** 0.5)
    assert torch.equal(got, dense_forward(x, wt)), "ordered partial sums: run to run identical"


@pytest.mark.parametrize("rows,dim", [(512, 96), (1000, 68), (8192, 256), (33, 4)])
def test_tower_training_column_passes_match_torch(rows, dim):
    """rf_column_stats / rf_activation_backward / rf_batchnorm_backward (the HBM-bound passes around a stage's three GEMMs)
    against torch float64, every activation whose derivative is taken from the output."""
    from recommendflow_b200 import dense_ops
    g = torch.Generator(device="cuda").manual_seed(rows + dim)
    x = torch.randn(rows, dim + 4, device="cuda", generator=g)[:, :dim] * 1.5 + 0.7           # a strided view: row pitch dim + 4
    mean, var, xt = dense_ops.column_stats(x, want_transpose=True)
    x64 = x.double()
    np.testing.assert_allclose(mean.cpu().numpy(), x64.mean(0).cpu().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(var.cpu().numpy(), x64.var(0, unbiased=False).cpu().numpy(), rtol=2e-5, atol=1e-6)
    assert torch.equal(xt, x.t().contiguous())
    assert torch.equal(mean, dense_ops.column_stats(x)[0]), "ordered partial sums: deterministic"
    dy = torch.randn(rows, dim, device="cuda", generator=g)
    for act, fwd in (("relu", torch.relu), ("selu", torch.selu), ("tanh", torch.tanh), ("sigmoid", torch.sigmoid), (None, lambda t: t)):
        z = torch.randn(rows, dim, device="cuda", generator=g).double().requires_grad_(True)
        y64 = fwd(z)
        (y64 * dy.double()).sum().backward()
        dz, dzt, db = dense_ops.activation_backward(dy, y64.detach().float().contiguous(), act)
        np.testing.assert_allclose(dz.cpu().numpy(), z.grad.cpu().numpy(), rtol=1e-4, atol=1e-5, err_msg=str(act))
        assert torch.equal(dzt, dz.t().contiguous())
        np.testing.assert_allclose(db.cpu().numpy(), z.grad.sum(0).cpu().numpy(), rtol=1e-4, atol=1e-4 * rows ** 0.5, err_msg=str(act))
    # BatchNormalization (batch statistics) backward vs autograd
    gamma = torch.rand(dim, device="cuda", generator=g) + 0.5
    eps = 1e-3
    xr = x64.clone().requires_grad_(True)
    g64 = gamma.double().requires_grad_(True)
    b64 = torch.zeros(dim, device="cuda", dtype=torch.float64, requires_grad=True)
    v64, m64 = torch.var_mean(xr, dim=0, unbiased=False)
    out = (xr - m64) * torch.rsqrt(v64 + eps) * g64 + b64
    dxh = torch.randn(rows, dim, device="cuda", generator=g)
    (out * dxh.double()).sum().backward()
    rstd = torch.rsqrt(var + eps)
    dx, dgamma, dbeta = dense_ops.batchnorm_backward(dxh, x, mean, rstd, gamma * rstd)
    scale = float(xr.grad.abs().max())
    np.testing.assert_allclose(dx.cpu().numpy(), xr.grad.cpu().numpy(), rtol=1e-3, atol=2e-5 * max(1.0, scale))
    np.testing.assert_allclose(dgamma.cpu().numpy(), g64.grad.cpu().numpy(), rtol=1e-3, atol=2e-4 * rows ** 0.5)
    np.testing.assert_allclose(dbeta.cpu().numpy(), b64.grad.cpu().numpy(), rtol=1e-3, atol=2e-4 * rows ** 0.5)


@pytest.mark.parametrize("activation,with_norm", [("selu", True), ("sigmoid", True), ("tanh", False), (None, True)])
def test_tower_mlp_training_stage_gradients(activation, with_norm):
    """The gradient-recording tower path (mlp._TrainStage: batch statistics folded into the tcgen05 GEMM, both backward
    GEMMs on the same kernel, BatchNormalization backward in closed form) against the same layers run one torch op at a
    time in float64 (Keras training=True semantics: biased batch variance, moving-average update)."""
    import copy
    from recommendflow_b200 import dense_ops
    from recommendflow_b200.backend.blocks.mlp import BatchNormalization, create_mlp
    torch.manual_seed(5)
    B, d_in, units = 512, 96, [128, 64]
    mlp = create_mlp(units, 0.0, activation, BatchNormalization(epsilon=1e-3) if with_norm else None, name="tower")
    x = (torch.randn(B, d_in, device="cuda") * 0.7 + 0.2)
    with torch.no_grad():
        mlp(x)                                        # creates the variables
    for p in mlp.parameters():
        p.requires_grad_(True)
        with torch.no_grad():
            if p.dim() == 1:
                p.add_(torch.rand_like(p) * 0.3)
    ref = copy.deepcopy(mlp).double()
    for m in list(mlp.modules()) + list(ref.modules()):
        if isinstance(m, BatchNormalization):
            m.batch_stats = True
    w = torch.randn(B, units[-1], device="cuda")
    xg = x.clone().requires_grad_(True)
    before = nat.launch_count()
    out = mlp(xg)
    (out * w).sum().backward()
    # per stage: forward = column statistics (2 launches, with a norm) + GEMM; backward = activation pass (2) + two GEMMs
    # + BatchNormalization backward (3, with a norm); the dW GEMM (K = batch) may add its split-K summation launch
    base = (10 if with_norm else 5) * len(units)
    assert base <= nat.launch_count() - before <= base + len(units)
    x64 = x.double().requires_grad_(True)
    old = dense_ops.DEFAULT_PRECISION
    dense_ops.DEFAULT_PRECISION = "fp32"              # the layer-by-layer path
    normed = []
    try:
        h = x64
        for layer in ref.layers:
            h = layer(h)
            if isinstance(layer, BatchNormalization):
                h.retain_grad()
                normed.append(h)
    finally:
        dense_ops.DEFAULT_PRECISION = old
    (h * w.double()).sum().backward()
    kinked = activation in ("selu", "relu")

    def close(a, b, what, atol=None):
        a, b = a.detach().double().cpu().numpy(), b.detach().cpu().numpy()
        rtol, atol = 2e-2, (6e-3 * max(1e-6, float(np.abs(b).max())) if atol is None else atol)
        if not kinked:
            np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, err_msg=what)
            return
        # selu' and relu' jump at z = 0 (selu: 1.758 -> 1.051): a pre-activation within TF32 rounding (~1e-3) of zero lands on
        # the other side of the jump than in float64 and changes that sample's whole dX row.  ~0.1 % of the B x units
        # pre-activations are that close, so a few percent of the rows differ (relu' jumps 0 -> 1, the largest effect); the
        # smooth cases of this test are held to the element-wise tolerance.
        ok = np.abs(a - b) <= atol + rtol * np.abs(b)
        assert ok.mean() >= 0.7, (what, float(ok.mean()))
        assert float(np.abs(a - b).mean()) <= max(atol, 5e-3 * float(np.abs(b).max())), what

    close(out, h, "forward")
    close(xg.grad, x64.grad, "dx")
    got, want = dict(mlp.named_parameters()), dict(ref.named_parameters())
    assert set(got) == set(want)
    # d gamma / d beta are column sums over the batch of dXhat (* xn): the next stage's normalisation makes the loss (nearly)
    # invariant to them, so the true sums cancel to ~0 while every summand carries TF32 rounding (2^-11 relative): the
    # tolerance is set from the size of the summands, sqrt(sum_i dXhat_i^2), not from the size of the result
    noise = {int(t.shape[1]): float(t.grad.pow(2).sum(dim=0).sqrt().max()) for t in normed}
    # a bias in front of a normalisation has a gradient of exactly 0 (the batch mean removes it); what the kernels return there
    # is fp32 summation noise, so every comparison also gets an absolute floor tied to the largest gradient of the stage stack
    floor = 2e-3 * max(float(p.grad.abs().max()) for p in want.values())
    for name in got:
        assert got[name].grad is not None, name
        atol = max(floor, 6e-3 * float(want[name].grad.abs().max()))
        if "gamma_" in name or "beta_" in name:
            atol = max(atol, (6e-2 if kinked else 2e-2) * noise[int(name.rsplit("_", 1)[1])])
        close(got[name].grad, want[name].grad, name, atol)
    for (n1, b1), (n2, b2) in zip(mlp.named_buffers(), ref.named_buffers()):        # moving statistics moved the same way
        close(b1, b2, n1)


def test_tower_mlp_fused_matches_layerwise_float64():
    """create_mlp([1024, 512, 256], 0.3, "selu", BatchNormalization(1e-6)) (dssm.py:25-26) at inference: the fused path
    (BN folded into each Dense, one tcgen05 launch per stage, l2 norm in the last epilogue) vs a float64 layer-by-layer
    restatement of the Keras semantics."""
    from recommendflow_b200 import _native as nat
    from recommendflow_b200.backend.blocks.mlp import BatchNormalization, create_mlp
    rng = np.random.default_rng(11)
    B, d_in, units = 2048, 1664, [1024, 512, 256]
    bn = BatchNormalization(epsilon=1e-6)
    mlp = create_mlp(units, 0.3, "selu", bn, name="tower")
    x = rng.standard_normal((B, d_in)).astype(np.float32) * 0.3
    dims, h64 = [d_in] + units, x.astype(np.float64)
    stage = 0
    for layer in mlp.layers:
        if hasattr(layer, "dense"):
            d0, d1 = dims[stage], dims[stage + 1]
            k = (rng.standard_normal((d0, d1)) / np.sqrt(d0)).astype(np.float32)
            b = rng.uniform(-0.1, 0.1, size=d1).astype(np.float32)
            layer.dense.set_weights([k, b])
            gamma, beta = rng.uniform(0.5, 1.5, d0).astype(np.float32), rng.uniform(-0.2, 0.2, d0).astype(np.float32)
            mean, var = rng.uniform(-0.3, 0.3, d0).astype(np.float32), rng.uniform(0.5, 2.0, d0).astype(np.float32)
            bn.set_weights([gamma, beta, mean, var])
            h64 = (h64 - mean) / np.sqrt(var.astype(np.float64) + 1e-6) * gamma + beta
            h64 = _act64("selu", h64 @ k.astype(np.float64) + b)
            stage += 1
    want = h64 / np.maximum(np.linalg.norm(h64, axis=1, keepdims=True), 1e-12)
    xt = torch.from_numpy(x).cuda()
    before = nat.launch_count()
    got = mlp(xt, l2_normalize=True)
    assert nat.launch_count() == before + 3, "three stages = three tensor-core launches"
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=2e-2, atol=2e-3)
    # every learned tensor is registered state
    keys = set(mlp.state_dict().keys())
    assert any("gamma_1664" in k for k in keys) and any("moving_var_512" in k for k in keys) and any("kernel" in k for k in keys)
    # a weight update invalidates the folded cache
    with torch.no_grad():
        mlp.layers[1].dense.bias.add_(1.0)
    got2 = mlp(xt, l2_normalize=True)
    assert not torch.allclose(got, got2)


def test_multi_head_attention_fused_qkv_path_matches_oracle():
    """Self-attention with one head at inference: one tcgen05 Dense produces q | k | v side by side and the tcgen05 SDPA
    kernel reads them as column windows (row pitch in the TMA descriptors).  vs the float64/fp32 oracle, TF32 tolerance."""
    from recommendflow_b200 import _native as nat
    rng = np.random.default_rng(31)
    B, S, d = 37, 50, 64
    x = rng.standard_normal((B, S, d)).astype(np.float32)
    lens = rng.integers(1, S + 1, size=B)
    mask = (np.arange(S)[None, :] < lens[:, None]).astype(np.float32)
    layer = MultiHeadAttention(d, 1)
    ws = []
    for dense in (layer.wq, layer.wk, layer.wv):
        w = (rng.standard_normal((d, d)) * 0.15).astype(np.float32)
        b = (rng.standard_normal(d) * 0.1).astype(np.float32)
        dense.set_weights([w, b])
        ws.append((w, b))
    xt = torch.from_numpy(x).cuda()
    before = nat.launch_count()
    got = layer(xt, xt, xt, torch.from_numpy(mask[:, :, None]).cuda())
    assert nat.launch_count() == before + 2, "one Dense launch + one SDPA launch"
    want = oracle.multi_head_attention(x, mask, *ws, 1)
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=2e-2, atol=2e-2)


def test_multi_head_attention_fused_qkv_gradients_match_float64():
    """The same layer under autograd: concatenated kernels on the tape, one differentiable tensor-core Dense
    (dense_ops.DenseFunction: forward, dX and the split-K dW on the tcgen05 kernel), and the attention backward writing
    dq | dk | dv as one buffer (rf_sdpa_backward_strided).  Against a float64 torch restatement of layer_utils.py:4-24."""
    from recommendflow_b200 import _native as nat
    torch.manual_seed(8)
    B, S, d = 40, 50, 64
    x = torch.randn(B, S, d, device="cuda")
    lens = torch.randint(1, S + 1, (B,), device="cuda")
    mask = (torch.arange(S, device="cuda")[None, :] < lens[:, None]).float()[:, :, None]
    layer = MultiHeadAttention(d, 1)
    with torch.no_grad():
        layer(x, x, x, mask)
    params = []
    for dense in (layer.wq, layer.wk, layer.wv):
        with torch.no_grad():
            dense.kernel.copy_(torch.randn(d, d, device="cuda") * 0.15)
            dense.bias.copy_(torch.randn(d, device="cuda") * 0.1)
        dense.kernel.requires_grad_(True)
        dense.bias.requires_grad_(True)
        params += [dense.kernel, dense.bias]
    w = torch.randn(B, S, d, device="cuda")
    xg = x.clone().requires_grad_(True)
    before = nat.launch_count()
    out = layer(xg, xg, xg, mask)
    (out * w).sum().backward()
    # forward: Dense + SDPA; backward: SDPA, column pass (2 launches), dX GEMM, dW GEMM (+ its split-K summation)
    assert 7 <= nat.launch_count() - before <= 8
    x64 = x.double().requires_grad_(True)
    p64 = [p.detach().double().requires_grad_(True) for p in params]
    q, k, v = (x64 @ p64[2 * i] + p64[2 * i + 1] for i in range(3))
    logits = q @ k.transpose(1, 2) / d ** 0.5
    logits = torch.where(mask.double().expand(-1, -1, S) == 0, torch.full_like(logits, -2.0 ** 32 + 1), logits)
    ref = torch.softmax(logits, dim=-1) @ v
    (ref * w.double()).sum().backward()
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref.detach().cpu().numpy(), rtol=2e-2, atol=2e-2)

    # the key bias has a gradient of exactly zero (softmax is invariant to a shift of a whole logits row): the absolute floor
    # comes from the largest parameter gradient, not from the compared tensor
    floor = 1e-3 * max(float(r.grad.abs().max()) for r in p64)

    def close(a, b, what):
        b = b.cpu().numpy()
        np.testing.assert_allclose(a.cpu().numpy(), b, rtol=2e-2, atol=max(floor, 5e-3 * float(np.abs(b).max())), err_msg=what)

    close(xg.grad, x64.grad, "dx")
    for i, (p, r) in enumerate(zip(params, p64)):
        close(p.grad, r.grad, f"param {i}")
