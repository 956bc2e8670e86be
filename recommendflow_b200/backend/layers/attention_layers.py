"""`MultiHeadAttention` and `SelfAttention` with the reference's constructor and call surface
(/root/reference/backend/layers/attention_layers.py:137-168 and :83-134).

The attention core runs in rf_sdpa_forward[_tc]; the Dense q/k/v projections run in rf_dense_forward_tc (the
tcgen05 GEMM with the bias / activation epilogue) whenever no gradient is being recorded -- under autograd (the
training step) they are library GEMMs (cuBLAS through torch.matmul), whose backward torch provides.  Keras `Dense`
defaults are kept: glorot-uniform kernel, zero bias, no activation, no output projection.
"""
import math

import numpy as np
import torch

from ... import dense_ops
from .layer_utils import scaled_dot_product_attention, split_heads
from .preprocess_layers import Layer


class _Tf32Flag(object):
    def __enter__(self):
        self.prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = dense_ops.DEFAULT_PRECISION == "tf32"

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32 = self.prev


class _LibraryMatmul(torch.autograd.Function):
    """x [..., in] @ w [in, out] with the precision mode applied to the BACKWARD products as well (torch's own MmBackward
    runs after the forward's flag has been restored, i.e. on the fp32 SIMT kernels)."""

    @staticmethod
    def forward(ctx, x, w):
        ctx.save_for_backward(x, w)
        with _Tf32Flag():
            return torch.matmul(x, w)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        with _Tf32Flag():
            gx = torch.matmul(g, w.t()) if ctx.needs_input_grad[0] else None
            gw = torch.matmul(x.reshape(-1, x.shape[-1]).t(), g.reshape(-1, g.shape[-1])) if ctx.needs_input_grad[1] else None
        return gx, gw


def library_matmul(x, w):
    """Plain library GEMM (cuBLAS).  In the default "tf32" mode fp32 operands go through the TF32 tensor-core
    path, which is also TensorFlow's default for fp32 matmuls on Ampere-and-later GPUs."""
    if w.dim() == 2 and torch.is_grad_enabled() and (x.requires_grad or w.requires_grad):
        return _LibraryMatmul.apply(x, w)
    with _Tf32Flag():
        return torch.matmul(x, w)


class Dense(Layer):
    """Keras Dense(units, activation=None): y = x W + b, W: [in, units] built on first call."""

    def __init__(self, units, activation=None, name=None):
        super().__init__(name=name)
        self.units, self.activation = units, activation
        self.kernel = self.bias = None

    def build(self, in_dim, device):
        if self.kernel is None:
            limit = math.sqrt(6.0 / (in_dim + self.units))          # glorot_uniform
            self.kernel = torch.nn.Parameter(torch.empty(in_dim, self.units, device=device).uniform_(-limit, limit),
                                             requires_grad=False)
            self.bias = torch.nn.Parameter(torch.zeros(self.units, device=device), requires_grad=False)
        return self

    def set_weights(self, weights):
        k, b = weights
        k = torch.as_tensor(np.asarray(k), dtype=torch.float32)
        dev = self.kernel.device if self.kernel is not None else torch.device("cuda", torch.cuda.current_device())
        self.kernel = torch.nn.Parameter(k.to(dev), requires_grad=False)
        self.bias = torch.nn.Parameter(torch.as_tensor(np.asarray(b), dtype=torch.float32).to(dev), requires_grad=False)
        self.units = k.shape[1]

    def get_weights(self):
        return [self.kernel.detach().cpu().numpy(), self.bias.detach().cpu().numpy()]

    def kernel_t(self):
        """[units, in] copy of the kernel (the K-major operand the tensor-core GEMM reads), refreshed when the
        kernel tensor is replaced or updated in place."""
        key = (self.kernel.data_ptr(), self.kernel._version)
        if getattr(self, "_kt_key", None) != key:
            self._kt = self.kernel.detach().t().contiguous()
            self._kt_key = key
        return self._kt

    def recording_grad(self, x):
        return torch.is_grad_enabled() and (x.requires_grad or self.kernel.requires_grad or self.bias.requires_grad)

    def call(self, x):
        self.build(x.shape[-1], x.device)
        if (dense_ops.DEFAULT_PRECISION == "tf32" and not self.recording_grad(x) and self.activation in (None, "relu")
                and dense_ops.dense_tc_ok(x, self.kernel.shape[0], self.units)):
            return dense_ops.dense_forward(x, self.kernel_t(), self.bias, self.activation)
        y = library_matmul(x, self.kernel) + self.bias
        return torch.relu(y) if self.activation == "relu" else y


class MultiHeadAttention(Layer):
    def __init__(self, d_model, num_heads):
        super().__init__(name="multi_head_attention")
        self.d_model = d_model
        self.num_heads = num_heads
        self.wq = Dense(d_model, activation=None)
        self.wk = Dense(d_model, activation=None)
        self.wv = Dense(d_model, activation=None)

    def _fused_qkv_weights(self):
        """[3 * d_model, in] weight and [3 * d_model] bias of ONE Dense producing q | k | v side by side."""
        key = tuple((d.kernel.data_ptr(), d.kernel._version, d.bias.data_ptr(), d.bias._version) for d in (self.wq, self.wk, self.wv))
        if getattr(self, "_qkv_key", None) != key:
            self._qkv_w = torch.cat([d.kernel.detach().t() for d in (self.wq, self.wk, self.wv)], dim=0).contiguous()
            self._qkv_b = torch.cat([d.bias.detach() for d in (self.wq, self.wk, self.wv)]).contiguous()
            self._qkv_key = key
        return self._qkv_w, self._qkv_b

    def call(self, q, k, v, mask):
        if (q is k and k is v and self.num_heads == 1 and dense_ops.DEFAULT_PRECISION == "tf32" and q.dim() == 3 and q.is_cuda
                and dense_ops.sdpa_tc_shape_ok(q.shape[1], self.d_model) and q.shape[-1] % 4 == 0):
            for d in (self.wq, self.wk, self.wv):
                d.build(q.shape[-1], q.device)
            if not any(d.recording_grad(q) for d in (self.wq, self.wk, self.wv)):
                # self-attention, one head, inference: ONE tensor-core Dense for the three projections (x is read once),
                # and the attention kernel reads q, k, v as column windows of its output
                w, b = self._fused_qkv_weights()
                qkv = dense_ops.dense_forward(q, w, b, None)
                return dense_ops.sdpa_fused_qkv(qkv, mask, self.d_model)
            if q.shape[0] * q.shape[1] % 4 == 0:
                # the same under autograd: the three kernels are concatenated on the tape (the gradient splits back to wq / wk /
                # wv), one differentiable tensor-core Dense, and the attention backward returns dq | dk | dv as its one dY
                wf = torch.cat([d.kernel for d in (self.wq, self.wk, self.wv)], dim=1)
                bf = torch.cat([d.bias for d in (self.wq, self.wk, self.wv)])
                qkv = dense_ops.dense_autograd(q, wf, bf, None)
                return dense_ops.sdpa_fused_qkv_autograd(qkv, mask, self.d_model)
        q, k, v = self.wq(q), self.wk(k), self.wv(v)                       # (B, S, d_model)
        seq_len, d_model = q.shape[1], q.shape[2]
        depth = d_model // self.num_heads
        q = split_heads(q, seq_len, self.num_heads, depth)                 # (B, H, S, depth)
        k = split_heads(k, seq_len, self.num_heads, depth)
        v = split_heads(v, seq_len, self.num_heads, depth)
        mask = mask.unsqueeze(1).expand(-1, self.num_heads, -1, -1)        # (B, H, S, 1)
        att = scaled_dot_product_attention(q, k, v, mask)                  # (B, H, S, depth)
        return att.permute(0, 2, 1, 3).reshape(-1, seq_len, d_model)       # no output projection (as the reference)

    def forward(self, q, k, v, mask):
        return self.call(q, k, v, mask)


class SelfAttention(Layer):
    """Shared-weight relu projections, sinusoidal positions, key... QUERY-row mask, mean over the sequence."""

    def __init__(self, add_pos=True):
        super().__init__(name="self_attention")
        self.add_pos = add_pos
        self.W = None

    def build(self, dim, device):
        if self.W is None:
            self.dim = dim
            self.W = torch.nn.Parameter(torch.empty(dim, dim, device=device).normal_(0.0, 0.05), requires_grad=False)
        return self

    @staticmethod
    def get_angles(pos, i, d_model):
        return pos * (1 / np.power(10000, (2 * (i // 2)) / np.float32(d_model)))

    def positional_encoding(self, qk):
        ang = self.get_angles(np.arange(qk.shape[1])[:, np.newaxis], np.arange(self.dim)[np.newaxis, :], self.dim)
        ang[:, 0::2] = np.sin(ang[:, 0::2])
        ang[:, 1::2] = np.cos(ang[:, 1::2])
        return torch.as_tensor(ang[np.newaxis, ...], dtype=torch.float32, device=qk.device)

    def call(self, inputs, **kwargs):
        q, k, v, mask = inputs
        self.build(q.shape[-1], q.device)
        if self.add_pos:
            k = k + self.positional_encoding(k)
            q = q + self.positional_encoding(q)
        q = torch.relu(library_matmul(q, self.W))
        k = torch.relu(library_matmul(k, self.W))
        # the reference scales by sqrt(self.dim) == sqrt(k.shape[-1]) and tiles the [B, S, 1] mask over keys
        out = scaled_dot_product_attention(q, k, v, mask)
        return out.mean(dim=1)
