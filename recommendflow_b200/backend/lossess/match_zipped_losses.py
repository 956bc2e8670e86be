"""Keras-style (y_true, y_pred) wrappers: rows of y_pred interleave query and doc embeddings
(/root/reference/backend/lossess/match_zipped_losses.py:7-28).  The reference casts to float64 after
l2-normalising; here the contraction runs in fp32 (documented tolerance in tests)."""
import torch

from . import match_losses


def zip_embedding(q, a):
    return torch.cat([q, a], dim=1).reshape(-1, a.shape[1])


def unzip_embedding(y_true, y_pred):
    y_true = y_true.squeeze(1) if y_true.dim() == 2 else y_true
    q = torch.nn.functional.normalize(y_pred[::2].to(torch.float32), dim=1, eps=1e-12)
    d = torch.nn.functional.normalize(y_pred[1::2].to(torch.float32), dim=1, eps=1e-12)
    return y_true.to(torch.float32), q, d


def _wrap(name):
    core = getattr(match_losses, name)

    def loss(y_true, y_pred, *args, **kwargs):
        y, q, d = unzip_embedding(y_true, y_pred)
        return core(y, q, d, *args, **kwargs)

    loss.__name__ = name
    loss.__doc__ = f"zipped form of match_losses.{name}"
    return loss


for _n in ("mean_squared_error", "binary_cross_entropy", "cosent_loss", "cosent_loss_v2", "batch_neg_sample_ce_loss",
           "batch_neg_sample_symmetrical_ce_loss", "batch_neg_sample_scaled_multi_class_ce_loss",
           "batch_neg_sample_symmetrical_scaled_multi_class_ce_loss", "batch_neg_sample_margin_rank_loss",
           "batch_hard_neg_sample_margin_rank_loss"):
    globals()[_n] = _wrap(_n)
