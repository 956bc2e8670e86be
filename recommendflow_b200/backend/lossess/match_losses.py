"""Two-tower losses with the reference's names and signatures
(/root/reference/backend/lossess/match_losses.py).

Every in-batch loss there starts from `tf.matmul(query, tf.transpose(doc))` -- a B x B matrix that
the reference materialises (268 MB at B = 8192; 17 GB at 65536) and then reduces row by row.  Here
the contraction and the row reductions happen in one pass inside rf_inbatch_rowstats (hand-written
CUDA, include/rf_b200.h); only [B]-sized vectors ever reach memory.  What remains in torch ops is
O(B) glue on those vectors.

Quirks kept from the reference: no max-subtraction is *visible* (we subtract the row max inside
the kernel, which is the same real number wherever the reference does not overflow); the
"symmetrical scaled" loss applies `scale` twice and reuses the row sums for its doc side
(:180-186); the margin-rank loss multiplies the B x B hinge matrix by a [B] y_true, i.e. weights
COLUMNS (:205).
"""

import torch

from ...dense_ops import inbatch_rowstats, inbatch_softmax_ce_autograd


def _vec(t, like):
    return torch.as_tensor(t, dtype=torch.float32, device=like.device).reshape(-1)


def _needs_grad(*tensors):
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)


class _RowDot(torch.autograd.Function):
    """diag(query . doc^T) = rowwise dot product: exact-fp32 CUDA forward (rf_inbatch_rowstats), elementwise backward."""

    @staticmethod
    def forward(ctx, query, doc):
        ctx.save_for_backward(query, doc)
        return inbatch_rowstats(query, doc, want=("diag",))["diag"]

    @staticmethod
    def backward(ctx, g):
        q, d = ctx.saved_tensors
        g = g.reshape(-1, 1)
        return (g * d) if ctx.needs_input_grad[0] else None, (g * q) if ctx.needs_input_grad[1] else None


def _diag(query, doc):
    """S_ii for every row; connected to the autograd graph when query / doc require grad."""
    if _needs_grad(query, doc):
        return _RowDot.apply(query.to(torch.float32), doc.to(torch.float32))
    return inbatch_rowstats(query, doc, want=("diag",))["diag"]


def mean_squared_error(y_true, query, doc):
    y_pred = _diag(query, doc)
    return torch.mean((_vec(y_true, y_pred) - y_pred) ** 2)


def binary_cross_entropy(y_true, query, doc):
    y_pred = _diag(query, doc)
    y = _vec(y_true, y_pred)
    p = torch.clamp(y_pred, 1e-7, 1 - 1e-7)                     # Keras epsilon clip
    return -(y * torch.log(p) + (1 - y) * torch.log(1 - p))


def _pairwise_logsumexp(s, y, positive_only):
    """log(1 + sum_{y_i < y_j [, s_i - s_j > 0]} exp(s_i - s_j)) in row chunks: each chunk is reduced to one
    logsumexp, so peak memory is O(chunk * B), never O(B * B)."""
    parts = [torch.zeros((), device=s.device, dtype=s.dtype)]          # the leading 0 term (:55, :68)
    for r0 in range(0, s.numel(), 2048):
        diff = s[r0:r0 + 2048, None] - s[None, :]
        keep = y[r0:r0 + 2048, None] < y[None, :]
        if positive_only:
            keep = keep & (diff > 0)
        parts.append(torch.logsumexp(torch.where(keep, diff, torch.full_like(diff, -1e12)).reshape(-1), dim=0))
    return torch.logsumexp(torch.stack(parts), dim=0)


def cosent_loss(y_true, query, doc, scale=20):
    """log(1 + sum_{y_i < y_j} exp(scale * (s_i - s_j))), s = rowwise query . doc   (:42-56).
    Not a q.d^T contraction: pairwise over the B scores, reduced chunk by chunk."""
    s = _diag(query, doc) * scale
    return _pairwise_logsumexp(s, _vec(y_true, s), False)


def cosent_loss_v2(y_true, query, doc, scale=20):
    """cosent with the negative differences dropped as well (:59-69)."""
    s = _diag(query, doc) * scale
    return _pairwise_logsumexp(s, _vec(y_true, s), True)


def _gather_rows(ind, *tensors):
    return [t[ind] for t in tensors]


def aux_label_cosent_loss(y_true, aux_true, query, doc, scale=20, alpha: float = .5):
    """(1 - alpha) * cosent_v2 over the positives' auxiliary labels + alpha * cosent_v2 over the negatives' (:72-96)."""
    q = torch.as_tensor(query)
    y = _vec(y_true, q)
    aux = _vec(aux_true, q)
    d = torch.as_tensor(doc, device=q.device)
    pos, neg = torch.nonzero(y == 1).reshape(-1), torch.nonzero(y == 0).reshape(-1)
    pos_loss = cosent_loss_v2(*_gather_rows(pos, aux, q, d), scale) if pos.numel() else torch.zeros((), device=q.device)
    neg_loss = cosent_loss_v2(*_gather_rows(neg, aux, q, d), scale) if neg.numel() else torch.zeros((), device=q.device)
    return (1 - alpha) * pos_loss + alpha * neg_loss


def pos_aux_label_cosent_loss(y_true, aux_true, query, doc, scale=20):
    """cosent_v2 of the auxiliary label over the positive samples only (:99-116)."""
    q = torch.as_tensor(query)
    y = _vec(y_true, q)
    pos = torch.nonzero(y == 1).reshape(-1)
    if not pos.numel():
        return torch.zeros((), device=q.device)
    return cosent_loss_v2(*_gather_rows(pos, _vec(aux_true, q), q, torch.as_tensor(doc, device=q.device)), scale)


def batch_neg_sample_scaled_multi_class_ce_loss(y_true, query, doc, scale=20):
    """mean_i( -log( exp(s S_ii) / sum_j exp(s S_ij) ) * y_i ),  S = query . doc^T   (:150-165).
    When query or doc requires grad the differentiable op (CUDA forward + CUDA backward) is recorded."""
    if _needs_grad(query, doc):
        return inbatch_softmax_ce_autograd(_vec(y_true, query), query, doc, float(scale))
    return inbatch_rowstats(query, doc, y_true=y_true, scale=scale, want=("lse", "diag"))["loss"]


def batch_neg_sample_symmetrical_scaled_multi_class_ce_loss(y_true, query, doc, scale=20):
    """Reference quirk: y_pred = scale * S is scaled AGAIN inside both exponentials and the "doc side"
    reuses the row sums, so the value equals the one-sided loss at temperature scale**2 (:169-189)."""
    s2 = float(scale) * float(scale)
    if _needs_grad(query, doc):
        return inbatch_softmax_ce_autograd(_vec(y_true, query), query, doc, s2)
    return inbatch_rowstats(query, doc, y_true=y_true, scale=s2, want=("lse", "diag"))["loss"]


class _RankLoss(torch.autograd.Function):
    """The two margin-rank losses as differentiable ops.  Forward: the fused row-statistics kernel (S never
    stored).  Backward: S is re-formed 1024 rows at a time (library GEMM), turned into the sparse 0/1 coefficient
    matrix of the hinge and contracted back -- O(chunk * B) memory.  These sibling losses are off the measured
    path (base_recall_sdpa trains with the scaled multi-class CE), hence no dedicated kernel."""

    @staticmethod
    def forward(ctx, y, query, doc, margin, hard):
        ctx.save_for_backward(y, query, doc)
        ctx.margin, ctx.hard = float(margin), bool(hard)
        if hard:
            r = inbatch_rowstats(query, doc, want=("diag", "maxoff"))
            return (torch.clamp(r["maxoff"] - r["diag"] + margin, 0, 1e14) * y).sum()
        return inbatch_rowstats(query, doc, col_weight=y, margin=margin, want=("hinge",))["hinge"].sum()

    @staticmethod
    def backward(ctx, g):
        y, q, d = ctx.saved_tensors
        B = q.shape[0]
        gq, gd = torch.zeros_like(q), torch.zeros_like(d)
        diag = (q * d).sum(dim=1)
        eye_cols = torch.arange(B, device=q.device)
        for r0 in range(0, B, 1024):
            r1 = min(B, r0 + 1024)
            S = q[r0:r1] @ d.t()                                           # [chunk, B]
            rows = torch.arange(r0, r1, device=q.device)
            on_diag = rows[:, None] == eye_cols[None, :]
            if ctx.hard:
                Sz = torch.where(on_diag, torch.zeros_like(S), S)          # diagonal zeroed (:219)
                mx, arg = Sz.max(dim=1)
                h = mx - diag[r0:r1] + ctx.margin
                w = ((h > 0) & (h < 1e14)).to(S.dtype) * y[r0:r1]          # d loss / d (max_i - S_ii)
                C = torch.zeros_like(S)
                C[torch.arange(r1 - r0, device=q.device), arg] = w
                C = torch.where(on_diag, torch.zeros_like(C), C)           # a zeroed diagonal entry carries no gradient
                C[torch.arange(r1 - r0, device=q.device), rows] -= w      # - S_ii
            else:
                h = S - diag[r0:r1, None] + ctx.margin
                C = ((h > 0) & (h < 1e14)).to(S.dtype) * y[None, :]        # d loss / d S_ij (column weights, :205)
                C = torch.where(on_diag, torch.zeros_like(C), C)           # S_ii - S_ii: constant
                C[torch.arange(r1 - r0, device=q.device), rows] = -C.sum(dim=1)   # - S_ii of every active term
            gq[r0:r1] = C @ d
            gd += C.t() @ q[r0:r1]
        return None, gq * g, gd * g, None, None


def batch_neg_sample_margin_rank_loss(y_true, query, doc, margin=0.1):
    """sum_ij clip(S_ij - S_ii + margin, 0, 1e14) * y_j   (:193-206; y_true [B] broadcasts over columns)."""
    if _needs_grad(query, doc):
        return _RankLoss.apply(_vec(y_true, query), query.to(torch.float32), doc.to(torch.float32), margin, False)
    r = inbatch_rowstats(query, doc, col_weight=y_true, margin=margin, want=("hinge",))
    return r["hinge"].sum()


def batch_hard_neg_sample_margin_rank_loss(y_true, query, doc, margin=0.1):
    """sum_i clip(max_j(S_ij with the diagonal zeroed) - S_ii + margin, 0, 1e14) * y_i   (:209-226)."""
    if _needs_grad(query, doc):
        return _RankLoss.apply(_vec(y_true, query), query.to(torch.float32), doc.to(torch.float32), margin, True)
    r = inbatch_rowstats(query, doc, want=("diag", "maxoff"))
    y = _vec(y_true, r["diag"])
    return (torch.clamp(r["maxoff"] - r["diag"] + margin, 0, 1e14) * y).sum()


def batch_neg_sample_ce_loss(y_true, query, doc):
    """K.categorical_crossentropy(diag(y), S) * y, mean   (:119-130).  Keras normalises the raw scores by
    their row sum and clips to [1e-7, 1 - 1e-7]; row sums of S are q . (sum_j d_j), a GEMV."""
    q = torch.as_tensor(query, dtype=torch.float32)
    d = torch.as_tensor(doc, dtype=torch.float32, device=q.device)
    diag = _diag(q, d)
    y = _vec(y_true, diag)
    rowsum = q @ d.sum(dim=0)
    p = torch.clamp(diag / rowsum, 1e-7, 1 - 1e-7)
    return torch.mean(-y * torch.log(p) * y)


def batch_neg_sample_symmetrical_ce_loss(y_true, query, doc):
    q = torch.as_tensor(query, dtype=torch.float32)
    d = torch.as_tensor(doc, dtype=torch.float32, device=q.device)
    diag = _diag(q, d)
    y = _vec(y_true, diag)
    p1 = torch.clamp(diag / (q @ d.sum(dim=0)), 1e-7, 1 - 1e-7)
    p2 = torch.clamp(diag / (d @ q.sum(dim=0)), 1e-7, 1 - 1e-7)
    return torch.mean(0.5 * (-y * torch.log(p1) - y * torch.log(p2)) * y)


def batch_softmax_probabilistic_combining_soft(batch_size, miu=0.6):
    raise NotImplementedError("batch_spc_soft needs thresholded pseudo-positive sums that are not part of the "
                              "row statistics kernel yet")


__all__ = [n for n in dir() if not n.startswith("_") and n not in ("torch", "inbatch_rowstats",
                                                                   "inbatch_softmax_ce_autograd")]
