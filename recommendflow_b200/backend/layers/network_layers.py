"""`TransformerEncoder` and `FFN` with the reference's surface
(/root/reference/backend/layers/network_layers.py:301-352).

Reference quirk kept: `TransformerEncoder.__init__` builds `tf.keras.layers.MultiHeadAttention(d_model,
num_heads)` (:331) although Keras' signature is `(num_heads, key_dim)` -- so the layer really has
`d_model` heads of size `num_heads` (size 1 by default), with Keras' own q/k/v/output projections.
`call([x, mask])` passes the [B, S, 1] mask as `attention_mask`, which broadcasts over KEYS: a zero
masks a whole QUERY row (Keras adds -1e9 to every logit of that row -> uniform attention), the same
effect as the reference's own `scaled_dot_product_attention`.

The attention core runs in rf_sdpa_forward (CUDA).  At inference the q / k / v / output projections and the 1x1-conv
FFN run on the tensor-core Dense kernel (rf_dense_forward_tc: the [in, N, H] einsum kernels are flattened to [in, N*H]);
under autograd they are library GEMMs (torch provides the backward).  LayerNorm is elementwise glue.  Dropout is
identity at inference (the reference default is dropout=0.).
"""
import math

import numpy as np
import torch

from ... import dense_ops

from .attention_layers import Dense
from .layer_utils import scaled_dot_product_attention
from .preprocess_layers import Layer


class LayerNormalization(Layer):
    def __init__(self, epsilon=1e-3, name=None):
        super().__init__(name=name)
        self.epsilon = epsilon
        self.gamma = self.beta = None

    def build(self, dim, device):
        if self.gamma is None:
            self.gamma = torch.nn.Parameter(torch.ones(dim, device=device), requires_grad=False)
            self.beta = torch.nn.Parameter(torch.zeros(dim, device=device), requires_grad=False)
        return self

    def call(self, x):
        self.build(x.shape[-1], x.device)
        return torch.nn.functional.layer_norm(x, (x.shape[-1],), self.gamma, self.beta, self.epsilon)


class KerasMultiHeadAttention(Layer):
    """tf.keras.layers.MultiHeadAttention(num_heads, key_dim): einsum projections to [B, S, N, H], scores
    scaled by 1/sqrt(H), additive -1e9 mask, softmax over keys, output projection [N, H] -> query dim."""

    def __init__(self, num_heads, key_dim, name=None):
        super().__init__(name=name)
        self.num_heads, self.key_dim = num_heads, key_dim
        self.wq = self.wk = self.wv = self.wo = None

    def build(self, dim, device):
        if self.wq is None:
            N, H = self.num_heads, self.key_dim
            lim = math.sqrt(6.0 / (dim + N * H))                  # glorot_uniform
            mk = lambda *shape: torch.nn.Parameter(torch.empty(*shape, device=device).uniform_(-lim, lim), requires_grad=False)
            zeros = lambda *shape: torch.nn.Parameter(torch.zeros(*shape, device=device), requires_grad=False)
            self.wq, self.bq = mk(dim, N, H), zeros(N, H)
            self.wk, self.bk = mk(dim, N, H), zeros(N, H)
            self.wv, self.bv = mk(dim, N, H), zeros(N, H)
            self.wo, self.bo = mk(N, H, dim), zeros(dim)
        return self

    def set_weights(self, weights):
        """[wq, bq, wk, bk, wv, bv, wo, bo] in Keras' variable order."""
        dev = torch.device("cuda", torch.cuda.current_device())
        names = ["wq", "bq", "wk", "bk", "wv", "bv", "wo", "bo"]
        for n, w in zip(names, weights):
            setattr(self, n, torch.nn.Parameter(torch.as_tensor(np.asarray(w), dtype=torch.float32).to(dev), requires_grad=False))

    def _project(self, x, w, b, tag):
        """einsum("abc,cde->abde", x, w) + b.  At inference the [in, N, H] kernel is flattened to one [in, N*H] Dense
        and runs on the tensor-core kernel (rf_dense_forward_tc); under autograd it stays the library einsum."""
        n_out = w.shape[1] * w.shape[2]
        recording = torch.is_grad_enabled() and (x.requires_grad or w.requires_grad or b.requires_grad)
        if (dense_ops.DEFAULT_PRECISION == "tf32" and not recording and dense_ops.dense_tc_ok(x, w.shape[0], n_out)):
            key = (w.data_ptr(), w._version)
            cache = self.__dict__.setdefault("_wt", {})
            if cache.get(tag, (None,))[0] != key:
                cache[tag] = (key, w.detach().reshape(w.shape[0], n_out).t().contiguous())
            y = dense_ops.dense_forward(x, cache[tag][1], b.detach().reshape(-1), None)
            return y.view(*x.shape[:-1], w.shape[1], w.shape[2])
        return torch.einsum("abc,cde->abde", x, w) + b

    def call(self, query, value, key=None, attention_mask=None):
        key = value if key is None else key
        self.build(query.shape[-1], query.device)
        B, T, _ = query.shape
        S = key.shape[1]
        N, H = self.num_heads, self.key_dim
        q = self._project(query, self.wq, self.bq, "q").permute(0, 2, 1, 3).reshape(B * N, T, H)
        k = self._project(key, self.wk, self.bk, "k").permute(0, 2, 1, 3).reshape(B * N, S, H)
        v = self._project(value, self.wv, self.bv, "v").permute(0, 2, 1, 3).reshape(B * N, S, H)
        mask = None
        if attention_mask is not None:
            if attention_mask.dim() != 3 or attention_mask.shape[-1] != 1:
                raise NotImplementedError("only the reference's [B, S, 1] query-row mask is supported")
            mask = attention_mask[:, None, :, :].expand(B, N, T, 1).reshape(B * N, T, 1)
        if T != S:
            raise NotImplementedError("self-attention only (the reference calls mha(x, x, x, mask))")
        ctx = scaled_dot_product_attention(q, k, v, mask)                       # [B*N, T, H]
        ctx = ctx.reshape(B, N, T, H).permute(0, 2, 1, 3)                       # [B, T, N, H]
        flat = ctx.reshape(B, T, N * H)
        recording = torch.is_grad_enabled() and (flat.requires_grad or self.wo.requires_grad or self.bo.requires_grad)
        if dense_ops.DEFAULT_PRECISION == "tf32" and not recording and dense_ops.dense_tc_ok(flat, N * H, self.wo.shape[-1]):
            key = (self.wo.data_ptr(), self.wo._version)
            cache = self.__dict__.setdefault("_wt", {})
            if cache.get("o", (None,))[0] != key:
                cache["o"] = (key, self.wo.detach().reshape(N * H, -1).t().contiguous())
            return dense_ops.dense_forward(flat.contiguous(), cache["o"][1], self.bo.detach(), None)
        return torch.einsum("abcd,cde->abe", ctx, self.wo) + self.bo


class FFN(Layer):
    """Conv1D(hidden, 1, relu) -> Conv1D(d_model, 1): two position-wise Dense layers."""

    def __init__(self, hidden_unit, d_model):
        super().__init__(name="ffn")
        self.conv1 = Dense(hidden_unit, activation="relu")
        self.conv2 = Dense(d_model, activation=None)

    def call(self, inputs):
        return self.conv2(self.conv1(inputs))


class TransformerEncoder(Layer):
    def __init__(self, d_model, num_heads=1, ffn_hidden_unit=128, dropout=0., layer_norm_eps=1e-6):
        super().__init__(name="transformer_encoder")
        self.mha = KerasMultiHeadAttention(d_model, num_heads)      # sic: (num_heads=d_model, key_dim=num_heads)
        self.ffn = FFN(ffn_hidden_unit, d_model)
        self.layernorm1 = LayerNormalization(epsilon=layer_norm_eps)
        self.layernorm2 = LayerNormalization(epsilon=layer_norm_eps)
        self.dropout = dropout

    def call(self, inputs):
        x, mask = inputs
        att_out = self.mha(x, x, x, mask)
        out1 = self.layernorm1(x + att_out)
        ffn_out = self.ffn(out1)
        return self.layernorm2(out1 + ffn_out)
