"""Host-side launchers of the dense contractions (rf_sdpa_forward, rf_inbatch_rowstats)."""
import ctypes as C
import os

import torch

from . import _native as nat


# "tf32": tcgen05 tensor cores, fp32 operands read as TF32 (what TensorFlow does for fp32 matmuls on
# Ampere and later GPUs); "fp32": exact CUDA-core accumulation.  Shapes the tensor-core kernels do
# not take use the fp32 kernels.
DEFAULT_PRECISION = "tf32"
TC_PRECISIONS = ("tf32", "bf16")   # operand formats the tensor-core kernels take (bf16: the in-batch logits only)
# The B x B in-batch logits default to bf16 operands in tensor-core mode: the TF32 kernel is bound by operand traffic
# (232 TFLOP/s at B = 8192, 319 at 65536), the bf16 kernel reaches 508 / 1121 TFLOP/s; the diagonal (the positives)
# stays an exact fp32 dot product and the tolerance is stated in tests/test_dense_gpu.py.  "tf32" / "fp32" on request.
DEFAULT_LOSS_PRECISION = "bf16"


def _f32(t, what):
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)
    if not t.is_cuda:
        raise nat.NativeError(f"{what} must live on a CUDA device (got {t.device}); there is no CPU fallback")
    if t.dtype == torch.float32 and t.is_contiguous():          # the common case: no dispatcher round trips
        return t
    return t.to(torch.float32).contiguous()


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


_on = nat.on_device


def dense_tc_ok(x, in_dim, units, l2_normalize=False):
    """Shapes / layouts rf_dense_forward_tc takes (everything else goes through the library GEMM)."""
    return (x.is_cuda and x.dtype == torch.float32 and in_dim % 4 == 0 and units % 4 == 0 and x.stride(-1) == 1
            and (not l2_normalize or units <= 256))


def dense_forward(x, weight_t, bias=None, activation=None, l2_normalize=False, out=None):
    """Keras Dense on the tensor cores: activation(x @ weight_t.T + bias) (then optionally l2-normalised rows).

    x: [..., in_dim] fp32 CUDA (rows may be strided: any view whose last dim is contiguous and whose leading dims
    collapse to one row pitch); weight_t: [units, in_dim] = the transposed Keras kernel; bias [units] or None.
    One launch of rf_dense_forward_tc (tcgen05, TF32 operands, fp32 accumulate, fused epilogue)."""
    if activation not in nat.ACTIVATION:
        raise ValueError(f"Unknown activation function: {activation}")
    w = _f32(weight_t, "weight_t")
    units, in_dim = w.shape
    if x.shape[-1] != in_dim:
        raise ValueError(f"input has {x.shape[-1]} features, the layer expects {in_dim}")
    lead = x.shape[:-1]
    x2 = x.reshape(-1, in_dim) if x.dim() != 2 else x
    if x2.dtype != torch.float32 or not x2.is_cuda:
        x2 = _f32(x2, "x")
    if x2.stride(1) != 1 or (x2.shape[0] > 1 and x2.stride(0) % 4) or x2.data_ptr() % 16:
        x2 = x2.contiguous()
    rows = x2.shape[0]
    ldx = x2.stride(0) if rows > 1 else in_dim
    if out is None:
        out2 = torch.empty(rows, units, dtype=torch.float32, device=x2.device)
    else:
        out2 = out.reshape(-1, units) if out.dim() != 2 else out
        if out2.dtype != torch.float32 or out2.shape[0] != rows or out2.stride(1) != 1:
            raise ValueError("out must be an fp32 [rows, units] view with contiguous columns")
    b = None if bias is None else _f32(bias, "bias")
    ws, need = None, 0
    if b is None and activation in (None, "linear") and not l2_normalize:        # a plain product may split K (few output tiles, long K)
        need = int(nat.lib().rf_dense_tc_workspace_bytes(rows, in_dim, units))
        if need:
            ws = torch.empty(need, dtype=torch.uint8, device=x2.device)
    with _on(x2.device):
        nat.check(nat.lib().rf_dense_forward_tc_ex(x2.data_ptr(), rows, in_dim, ldx, w.data_ptr(), None if b is None else b.data_ptr(),
                                                   units, nat.ACTIVATION[activation], 1 if l2_normalize else 0, out2.data_ptr(),
                                                   out2.stride(0) if rows > 1 else units, None if ws is None else ws.data_ptr(), need,
                                                   _stream(x2.device)))
    return out2.view(*lead, units) if out is None else out


def sdpa_tc_shape_ok(S, dh):
    return 1 <= S <= 64 and dh in (32, 64, 96)


def _tower_ws(rows, dim, device):
    need = int(nat.lib().rf_tower_train_workspace_bytes(rows, dim))
    return torch.empty(need, dtype=torch.uint8, device=device), need


def column_stats(x, want_transpose=False):
    """(mean, biased variance[, x^T]) of the columns of x [rows, dim] in one pass (rf_column_stats)."""
    rows, dim = x.shape
    mean = torch.empty(dim, dtype=torch.float32, device=x.device)
    var = torch.empty_like(mean)
    xt = torch.empty(dim, rows, dtype=torch.float32, device=x.device) if want_transpose else None
    ws, need = _tower_ws(rows, dim, x.device)
    with _on(x.device):
        nat.check(nat.lib().rf_column_stats(x.data_ptr(), rows, dim, x.stride(0), mean.data_ptr(), var.data_ptr(),
                                            None if xt is None else xt.data_ptr(), ws.data_ptr(), need, _stream(x.device)))
    return mean, var, xt


def activation_backward(dy, y, activation):
    """(dZ, dZ^T, db) from dY and the stage output y (rf_activation_backward).  Without an activation dZ is dY itself."""
    rows, units = dy.shape
    dy = dy if dy.is_contiguous() else dy.contiguous()
    plain = activation in (None, "linear")
    dz = dy if plain else torch.empty(rows, units, dtype=torch.float32, device=dy.device)
    dzt = torch.empty(units, rows, dtype=torch.float32, device=dy.device)
    db = torch.empty(units, dtype=torch.float32, device=dy.device)
    ws, need = _tower_ws(rows, units, dy.device)
    with _on(dy.device):
        nat.check(nat.lib().rf_activation_backward(dy.data_ptr(), None if plain else y.data_ptr(), rows, units, nat.ACTIVATION[activation],
                                                   None if plain else dz.data_ptr(), dzt.data_ptr(), db.data_ptr(), ws.data_ptr(), need,
                                                   _stream(dy.device)))
    return dz, dzt, db


def batchnorm_backward(dxh, x, mean, rstd, scale):
    """(dX, dgamma, dbeta) of BatchNormalization on batch statistics (rf_batchnorm_backward)."""
    rows, dim = x.shape
    dx = torch.empty(rows, dim, dtype=torch.float32, device=x.device)
    dgamma = torch.empty(dim, dtype=torch.float32, device=x.device)
    dbeta = torch.empty_like(dgamma)
    ws, need = _tower_ws(rows, dim, x.device)
    with _on(x.device):
        nat.check(nat.lib().rf_batchnorm_backward(dxh.data_ptr(), x.data_ptr(), x.stride(0), mean.data_ptr(), rstd.data_ptr(),
                                                  scale.data_ptr(), rows, dim, dgamma.data_ptr(), dbeta.data_ptr(), dx.data_ptr(),
                                                  ws.data_ptr(), need, _stream(x.device)))
    return dx, dgamma, dbeta


def sdpa(q, k, v, mask=None, precision=None):
    """q, k, v: [..., S, dh]; mask: [..., S, 1] (or [..., S]) of 0/1 -- the reference's query-row mask.
    precision "tf32" runs on the tensor cores when the shape allows (S <= 64, dh in 32/64/96)."""
    precision = precision or DEFAULT_PRECISION
    if precision not in ("tf32", "fp32"):
        raise ValueError("precision must be 'tf32' or 'fp32'")
    q, k, v = _f32(q, "q"), _f32(k, "k"), _f32(v, "v")
    if q.shape != k.shape or q.shape != v.shape:
        raise ValueError(f"q, k, v must share one shape, got {tuple(q.shape)}, {tuple(k.shape)}, {tuple(v.shape)}")
    S, dh = q.shape[-2], q.shape[-1]
    nb = q.numel() // (S * dh) if S * dh else 0
    m = None
    if mask is not None:
        m = _f32(mask, "mask")
        if m.dim() == q.dim() and m.shape[-1] == 1:
            m = m[..., 0]
        m = m.expand(q.shape[:-1]).contiguous()
    out = torch.empty_like(q)
    fn = nat.lib().rf_sdpa_forward_tc if (precision == "tf32" and sdpa_tc_shape_ok(S, dh)) else nat.lib().rf_sdpa_forward
    with _on(q.device):
        nat.check(fn(q.data_ptr(), k.data_ptr(), v.data_ptr(), None if m is None else m.data_ptr(),
                     nb, S, dh, out.data_ptr(), _stream(q.device)))
    return out


def sdpa_fused_qkv(qkv, mask, dh):
    """scaled_dot_product_attention on q, k, v that sit side by side in ONE [..., S, 3 * dh] projection output
    (rf_sdpa_forward_tc_strided: the TMA descriptors carry the row pitch; nothing is copied)."""
    qkv = _f32(qkv, "qkv")
    S = qkv.shape[-2]
    if qkv.shape[-1] != 3 * dh or not sdpa_tc_shape_ok(S, dh):
        raise ValueError("sdpa_fused_qkv takes [..., S, 3 * dh] with S <= 64 and dh in (32, 64, 96)")
    nb = qkv.numel() // (S * 3 * dh)
    flat = qkv.view(-1, 3 * dh)
    m = None
    if mask is not None:
        m = _f32(mask, "mask")
        if m.dim() == qkv.dim() and m.shape[-1] == 1:
            m = m[..., 0]
        m = m.expand(qkv.shape[:-1]).contiguous()
    out = torch.empty(*qkv.shape[:-1], dh, dtype=torch.float32, device=qkv.device)
    with _on(qkv.device):
        nat.check(nat.lib().rf_sdpa_forward_tc_strided(flat.data_ptr(), flat.data_ptr() + 4 * dh, flat.data_ptr() + 8 * dh, 3 * dh,
                                                       None if m is None else m.data_ptr(), nb, S, dh, out.data_ptr(),
                                                       _stream(qkv.device)))
    return out


def inbatch_rowstats(query, doc, y_true=None, col_weight=None, scale=20.0, margin=0.0, want=("lse", "diag"),
                     precision=None):
    """Row statistics of S = query . doc^T without materialising S.  Returns a dict with the
    requested [B] vectors among lse / diag / hinge / maxoff, plus "loss" when y_true is given."""
    precision = precision or (DEFAULT_LOSS_PRECISION if DEFAULT_PRECISION == "tf32" else DEFAULT_PRECISION)
    if precision not in ("tf32", "fp32", "bf16"):
        raise ValueError("precision must be 'tf32', 'bf16' or 'fp32'")
    q, d = _f32(query, "query"), _f32(doc, "doc")
    if q.dim() != 2 or q.shape != d.shape:
        raise ValueError(f"query and doc must both be [B, D], got {tuple(q.shape)} and {tuple(d.shape)}")
    B, D = q.shape
    dev = q.device
    y = None if y_true is None else _f32(y_true, "y_true").reshape(-1)
    cw = None if col_weight is None else _f32(col_weight, "col_weight").reshape(-1)
    if y is not None and y.numel() != B:
        raise ValueError("y_true must have one entry per row")
    use_bf16 = precision == "bf16" and D % 8 == 0 and D <= 256
    use_tc = use_bf16 or (precision in ("tf32", "bf16") and D % 4 == 0)
    ws_bytes = nat.lib().rf_inbatch_workspace_bytes_tc(B, D) if use_tc else nat.lib().rf_inbatch_workspace_bytes(B)
    ws = torch.empty(max(1, ws_bytes), dtype=torch.uint8, device=dev)
    res = {k: torch.empty(B, dtype=torch.float32, device=dev) for k in want}
    loss = torch.zeros((), dtype=torch.float32, device=dev) if y is not None else None
    ptr = lambda t: None if t is None else t.data_ptr()
    fn = nat.lib().rf_inbatch_rowstats_bf16 if use_bf16 else (nat.lib().rf_inbatch_rowstats_tc if use_tc else nat.lib().rf_inbatch_rowstats)
    with _on(dev):
        nat.check(fn(q.data_ptr(), d.data_ptr(), ptr(y), ptr(cw), B, D, float(scale), float(margin),
                     ws.data_ptr(), ptr(res.get("lse")), ptr(res.get("diag")), ptr(res.get("hinge")),
                     ptr(res.get("maxoff")), ptr(loss), _stream(dev)))
    if loss is not None:
        res["loss"] = loss
    return res


def _mask_rows(mask, q):
    if mask is None:
        return None
    m = _f32(mask, "mask")
    if m.dim() == q.dim() and m.shape[-1] == 1:
        m = m[..., 0]
    return m.expand(q.shape[:-1]).contiguous()


def sdpa_backward_tc_ok(S, dh):
    return S <= 64 and dh in (32, 64, 96, 128)


# The tensor-core backward (warp-level mma.sync, TF32) measured 0.749 ms against 0.732 ms for the exact-fp32 register-tiled kernel
# at [8192, 50, 64] on B200 (profiles/r2e_bench_train_c3_sdpa_bwd_tc.json): the legacy MMA path buys nothing there, so autograd
# uses the exact kernel; RF_SDPA_BWD_TC=1 selects the tensor-core one.
SDPA_BACKWARD_TC = os.environ.get("RF_SDPA_BWD_TC", "0") == "1"


def sdpa_backward(q, k, v, mask, grad_out, precision="fp32"):
    """(dq, dk, dv) of `sdpa` given grad_out = dL/d(out).  precision "fp32": rf_sdpa_backward (exact fp32, register-tiled CUDA
    cores); "tf32": rf_sdpa_backward_tc (the five products on the tensor cores) when the shape allows."""
    q, k, v, g = _f32(q, "q"), _f32(k, "k"), _f32(v, "v"), _f32(grad_out, "grad_out")
    if not (q.shape == k.shape == v.shape == g.shape):
        raise ValueError("q, k, v and grad_out must share one shape")
    S, dh = q.shape[-2], q.shape[-1]
    nb = q.numel() // (S * dh) if S * dh else 0
    m = _mask_rows(mask, q)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
    with _on(q.device):
        if precision != "fp32" and sdpa_backward_tc_ok(S, dh) and nb:
            nat.check(nat.lib().rf_sdpa_backward_tc(q.data_ptr(), k.data_ptr(), v.data_ptr(), dh, None if m is None else m.data_ptr(),
                                                    g.data_ptr(), nb, S, dh, dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), dh,
                                                    _stream(q.device)))
        else:
            nat.check(nat.lib().rf_sdpa_backward(q.data_ptr(), k.data_ptr(), v.data_ptr(), None if m is None else m.data_ptr(),
                                                 g.data_ptr(), nb, S, dh, dq.data_ptr(), dk.data_ptr(), dv.data_ptr(),
                                                 _stream(q.device)))
    return dq, dk, dv


def inbatch_softmax_ce_backward(query, doc, y_true, lse, scale=20.0, upstream=1.0, need_query=True, need_doc=True,
                                positives_on_diagonal=True, precision=None):
    """(d loss / d query, d loss / d doc) of batch_neg_sample_scaled_multi_class_ce_loss, from the forward's lse.
    positives_on_diagonal=False: `doc` is a block of negatives only (another rank's docs in the data-parallel step)
    and `lse` is the log-sum-exp over the whole row of the all-gathered logits."""
    q, d = _f32(query, "query"), _f32(doc, "doc")
    y, lse = _f32(y_true, "y_true").reshape(-1), _f32(lse, "lse").reshape(-1)
    B, D = q.shape
    if d.shape != q.shape or y.numel() != B or lse.numel() != B:
        raise ValueError("query / doc must be [B, D]; y_true and lse one entry per row")
    precision = precision or DEFAULT_PRECISION
    if precision != "fp32" and B % 4 == 0 and D % 4 == 0 and B >= 512:
        # tensor cores: the three contractions go through the tcgen05 Dense kernel over slabs of query rows
        gq, gd = torch.empty_like(q), torch.empty_like(d)
        need = int(nat.lib().rf_inbatch_ce_backward_tc_workspace_bytes(B, D))
        ws = torch.empty(need, dtype=torch.uint8, device=q.device)
        with _on(q.device):
            nat.check(nat.lib().rf_inbatch_softmax_ce_backward_tc(q.data_ptr(), d.data_ptr(), y.data_ptr(), lse.data_ptr(), B, D,
                                                                  float(scale), float(upstream), 1 if positives_on_diagonal else 0,
                                                                  ws.data_ptr(), need, gq.data_ptr(), gd.data_ptr(), _stream(q.device)))
        return (gq if need_query else None), (gd if need_doc else None)
    gq = torch.empty_like(q) if need_query else None
    gd = torch.empty_like(d) if need_doc else None
    with _on(q.device):
        nat.check(nat.lib().rf_inbatch_softmax_ce_backward_block(q.data_ptr(), d.data_ptr(), y.data_ptr(), lse.data_ptr(), B, D,
                                                                 float(scale), float(upstream), 1 if positives_on_diagonal else 0,
                                                                 None if gq is None else gq.data_ptr(),
                                                                 None if gd is None else gd.data_ptr(), _stream(q.device)))
    return gq, gd


class SdpaFunction(torch.autograd.Function):
    """scaled_dot_product_attention with the CUDA forward (tensor cores when the shape allows) and the CUDA
    backward, for training through torch.autograd."""

    @staticmethod
    def forward(ctx, q, k, v, mask, precision):
        ctx.save_for_backward(q, k, v, mask if mask is not None else torch.empty(0, device=q.device))
        ctx.has_mask = mask is not None
        ctx.precision = precision or DEFAULT_PRECISION
        return sdpa(q, k, v, mask, precision)

    @staticmethod
    def backward(ctx, grad_out):
        q, k, v, mask = ctx.saved_tensors
        dq, dk, dv = sdpa_backward(q, k, v, mask if ctx.has_mask else None, grad_out, ctx.precision if SDPA_BACKWARD_TC else "fp32")
        return dq.view_as(q), dk.view_as(k), dv.view_as(v), None, None


class InbatchSoftmaxCeFunction(torch.autograd.Function):
    """batch_neg_sample_scaled_multi_class_ce_loss(y_true, query, doc, scale) as one differentiable op: the
    forward keeps only the per-row log-sum-exp, the backward recomputes the logits tile by tile."""

    @staticmethod
    def forward(ctx, y_true, query, doc, scale, precision):
        res = inbatch_rowstats(query, doc, y_true=y_true, scale=scale, want=("lse",), precision=precision)
        ctx.save_for_backward(y_true, query, doc, res["lse"])
        ctx.scale, ctx.precision = float(scale), precision
        return res["loss"]

    @staticmethod
    def backward(ctx, grad_loss):
        y, q, d, lse = ctx.saved_tensors
        # upstream stays on the device (float(grad_loss) would stall the host on everything queued so far)
        gq, gd = inbatch_softmax_ce_backward(q, d, y, lse, ctx.scale, 1.0, ctx.needs_input_grad[1],
                                             ctx.needs_input_grad[2], precision="fp32" if ctx.precision == "fp32" else None)
        return None, (None if gq is None else gq.mul_(grad_loss)), (None if gd is None else gd.mul_(grad_loss)), None, None


class DenseFunction(torch.autograd.Function):
    """act(x W + b) with all three products on the tcgen05 Dense kernel: forward (fused bias + activation), dX = dZ W^T,
    dW = X^T dZ (split-K: the contraction runs over the rows); dZ^T and db come out of one pass over dY."""

    @staticmethod
    def forward(ctx, x, kernel, bias, activation):
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1])
        x2 = x2 if x2.stride(1) == 1 and x2.stride(0) % 4 == 0 and x2.data_ptr() % 16 == 0 else x2.contiguous()
        y = dense_forward(x2, kernel.t().contiguous(), bias, activation)
        ctx.activation, ctx.lead = activation, lead
        ctx.save_for_backward(x2, y, kernel)
        return y.view(*lead, kernel.shape[1])

    @staticmethod
    def backward(ctx, dy):
        x2, y, kernel = ctx.saved_tensors
        dz, dzt, db = activation_backward(dy.reshape(-1, dy.shape[-1]), y, ctx.activation)
        dx = dense_forward(dz, kernel, None, None).view(*ctx.lead, kernel.shape[0]) if ctx.needs_input_grad[0] else None
        dw = dense_forward(x2.t().contiguous(), dzt, None, None)
        return dx, dw, db, None


def dense_autograd(x, kernel, bias, activation=None):
    """Differentiable Keras Dense (kernel [in, units]) on the tensor cores; activation: None, relu, selu, tanh or sigmoid."""
    return DenseFunction.apply(x, kernel, bias, activation)


class SdpaFusedQkvFunction(torch.autograd.Function):
    """`sdpa_fused_qkv` with its backward: the gradient comes out as ONE [..., S, 3 * dh] buffer (dq | dk | dv side by side,
    rf_sdpa_backward_strided), i.e. directly the dY of the fused projection."""

    @staticmethod
    def forward(ctx, qkv, mask, dh):
        qkv = qkv.contiguous()
        ctx.dh = dh
        ctx.save_for_backward(qkv, mask)
        return sdpa_fused_qkv(qkv, mask, dh)

    @staticmethod
    def backward(ctx, grad_out):
        qkv, mask = ctx.saved_tensors
        dh, S = ctx.dh, qkv.shape[-2]
        g = _f32(grad_out, "grad_out").contiguous()
        nb = qkv.numel() // (S * 3 * dh)
        m = None
        if mask is not None:
            m = _f32(mask, "mask")
            if m.dim() == qkv.dim() and m.shape[-1] == 1:
                m = m[..., 0]
            m = m.expand(qkv.shape[:-1]).contiguous()
        dqkv = torch.empty_like(qkv)
        p, d = qkv.data_ptr(), dqkv.data_ptr()
        fn = (nat.lib().rf_sdpa_backward_tc if (SDPA_BACKWARD_TC and DEFAULT_PRECISION != "fp32" and sdpa_backward_tc_ok(S, dh))
              else nat.lib().rf_sdpa_backward_strided)
        with _on(qkv.device):
            nat.check(fn(p, p + 4 * dh, p + 8 * dh, 3 * dh, None if m is None else m.data_ptr(), g.data_ptr(),
                         nb, S, dh, d, d + 4 * dh, d + 8 * dh, 3 * dh, _stream(qkv.device)))
        return dqkv, None, None


def sdpa_fused_qkv_autograd(qkv, mask, dh):
    return SdpaFusedQkvFunction.apply(qkv, mask, dh)


def sdpa_autograd(q, k, v, mask=None, precision=None):
    return SdpaFunction.apply(q, k, v, mask, precision)


def inbatch_softmax_ce_autograd(y_true, query, doc, scale=20.0, precision=None):
    return InbatchSoftmaxCeFunction.apply(y_true, query, doc, scale, precision)
