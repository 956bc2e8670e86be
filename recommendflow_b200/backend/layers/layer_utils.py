"""Attention helpers with the reference's names (/root/reference/backend/layers/layer_utils.py)."""
import torch

from ...dense_ops import sdpa, sdpa_autograd


def scaled_dot_product_attention(q, k, v, mask):
    """softmax(where(mask == 0, -2**32 + 1, q k^T / sqrt(dk))) v   (layer_utils.py:4-24).

    q, k, v: [..., seq_len, dim] CUDA tensors; mask: [..., seq_len, 1] -- it broadcasts over KEYS, so a
    zero masks a whole QUERY row (which then attends uniformly), exactly as in the reference.
    When an input requires grad the differentiable op (CUDA forward + CUDA backward) is recorded."""
    if torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in (q, k, v)):
        return sdpa_autograd(q, k, v, mask)
    return sdpa(q, k, v, mask)


def split_heads(x, seq_len, num_heads, depth):
    """[B, seq_len, num_heads * depth] -> [B, num_heads, seq_len, depth]   (layer_utils.py:27-38)."""
    return x.reshape(-1, seq_len, num_heads, depth).permute(0, 2, 1, 3)


def index_mapping(inputs_dict, map_dict):
    """Feature index mapping (layer_utils.py:41-53)."""
    out = {}
    for key, value in inputs_dict.items():
        if map_dict.get(key) is None:
            raise ValueError("map dict error!")
        out[key] = (value + torch.as_tensor(map_dict[key])).reshape(-1, 1)
    return out
