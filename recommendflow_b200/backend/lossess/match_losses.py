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
from functools import partial

import torch

from ...dense_ops import inbatch_rowstats, inbatch_softmax_ce_autograd


def _vec(t, like):
    return torch.as_tensor(t, dtype=torch.float32, device=like.device).reshape(-1)


def mean_squared_error(y_true, query, doc):
    y_pred = inbatch_rowstats(query, doc, want=("diag",))["diag"]
    return torch.mean((_vec(y_true, y_pred) - y_pred) ** 2)


def binary_cross_entropy(y_true, query, doc):
    y_pred = inbatch_rowstats(query, doc, want=("diag",))["diag"]
    y = _vec(y_true, y_pred)
    p = torch.clamp(y_pred, 1e-7, 1 - 1e-7)                     # Keras epsilon clip
    return -(y * torch.log(p) + (1 - y) * torch.log(1 - p))


def cosent_loss(y_true, query, doc, scale=20):
    """log(1 + sum_{y_i < y_j} exp(scale * (s_i - s_j))), s = rowwise query . doc   (:42-56).
    Not a q.d^T contraction: pairwise over the B scores, done in chunks of rows."""
    s = inbatch_rowstats(query, doc, want=("diag",))["diag"] * scale
    y = _vec(y_true, s)
    parts = [torch.zeros(1, device=s.device)]
    for r0 in range(0, s.numel(), 4096):
        diff = s[r0:r0 + 4096, None] - s[None, :]
        keep = y[r0:r0 + 4096, None] < y[None, :]
        parts.append(torch.where(keep, diff, torch.full_like(diff, -1e12)).reshape(-1))
    return torch.logsumexp(torch.cat(parts), dim=0)


def cosent_loss_v2(y_true, query, doc, scale=20):
    s = inbatch_rowstats(query, doc, want=("diag",))["diag"] * scale
    y = _vec(y_true, s)
    parts = [torch.zeros(1, device=s.device)]
    for r0 in range(0, s.numel(), 4096):
        diff = s[r0:r0 + 4096, None] - s[None, :]
        keep = (y[r0:r0 + 4096, None] < y[None, :]) & (diff > 0)
        parts.append(torch.where(keep, diff, torch.full_like(diff, -1e12)).reshape(-1))
    return torch.logsumexp(torch.cat(parts), dim=0)


def batch_neg_sample_scaled_multi_class_ce_loss(y_true, query, doc, scale=20):
    """mean_i( -log( exp(s S_ii) / sum_j exp(s S_ij) ) * y_i ),  S = query . doc^T   (:150-165).
    When query or doc requires grad the differentiable op (CUDA forward + CUDA backward) is recorded."""
    if torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in (query, doc)):
        return inbatch_softmax_ce_autograd(_vec(y_true, query), query, doc, float(scale))
    return inbatch_rowstats(query, doc, y_true=y_true, scale=scale, want=("lse", "diag"))["loss"]


def batch_neg_sample_symmetrical_scaled_multi_class_ce_loss(y_true, query, doc, scale=20):
    """Reference quirk: y_pred = scale * S is scaled AGAIN inside both exponentials and the "doc side"
    reuses the row sums, so the value equals the one-sided loss at temperature scale**2 (:169-189)."""
    return inbatch_rowstats(query, doc, y_true=y_true, scale=float(scale) * float(scale), want=("lse", "diag"))["loss"]


def batch_neg_sample_margin_rank_loss(y_true, query, doc, margin=0.1):
    """sum_ij clip(S_ij - S_ii + margin, 0, 1e14) * y_j   (:193-206; y_true [B] broadcasts over columns)."""
    r = inbatch_rowstats(query, doc, col_weight=y_true, margin=margin, want=("hinge",))
    return r["hinge"].sum()


def batch_hard_neg_sample_margin_rank_loss(y_true, query, doc, margin=0.1):
    """sum_i clip(max_j(S_ij with the diagonal zeroed) - S_ii + margin, 0, 1e14) * y_i   (:209-226)."""
    r = inbatch_rowstats(query, doc, want=("diag", "maxoff"))
    y = _vec(y_true, r["diag"])
    return (torch.clamp(r["maxoff"] - r["diag"] + margin, 0, 1e14) * y).sum()


def batch_neg_sample_ce_loss(y_true, query, doc):
    """K.categorical_crossentropy(diag(y), S) * y, mean   (:119-130).  Keras normalises the raw scores by
    their row sum and clips to [1e-7, 1 - 1e-7]; row sums of S are q . (sum_j d_j), a GEMV."""
    q = torch.as_tensor(query, dtype=torch.float32)
    d = torch.as_tensor(doc, dtype=torch.float32, device=q.device)
    diag = inbatch_rowstats(q, d, want=("diag",))["diag"]
    y = _vec(y_true, diag)
    rowsum = q @ d.sum(dim=0)
    p = torch.clamp(diag / rowsum, 1e-7, 1 - 1e-7)
    return torch.mean(-y * torch.log(p) * y)


def batch_neg_sample_symmetrical_ce_loss(y_true, query, doc):
    q = torch.as_tensor(query, dtype=torch.float32)
    d = torch.as_tensor(doc, dtype=torch.float32, device=q.device)
    diag = inbatch_rowstats(q, d, want=("diag",))["diag"]
    y = _vec(y_true, diag)
    p1 = torch.clamp(diag / (q @ d.sum(dim=0)), 1e-7, 1 - 1e-7)
    p2 = torch.clamp(diag / (d @ q.sum(dim=0)), 1e-7, 1 - 1e-7)
    return torch.mean(0.5 * (-y * torch.log(p1) - y * torch.log(p2)) * y)


def batch_softmax_probabilistic_combining_soft(batch_size, miu=0.6):
    raise NotImplementedError("batch_spc_soft needs thresholded pseudo-positive sums that are not part of the "
                              "row statistics kernel yet")


__all__ = [n for n in dir() if not n.startswith("_") and n not in ("partial", "torch", "inbatch_rowstats")]
