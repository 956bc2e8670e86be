"""Tower MLP factory with the reference's signature (/root/reference/backend/blocks/mlp.py:4-15):
`create_mlp(hidden_units, dropout_rate, activation, normalization_layer)` -> [norm, Dense, Dropout] * n.

Reference quirk kept: the SAME normalization layer instance is appended before every Dense (:11), so
one BatchNormalization would be applied to inputs of different widths -- Keras builds it on the first
width and fails on the second unless all widths are equal.  Here the shared instance keeps one set of
statistics per distinct width (the only way the reference's towers [1024, 512, 256] can run at all).

Inference (no gradient being recorded, moving statistics, no dropout): every [norm, Dense, activation] stage is
ONE launch of the tcgen05 GEMM rf_dense_forward_tc -- BatchNormalization is an affine map per input column there
and is folded into the Dense kernel and bias (cached until a weight changes); the last stage can also l2-normalise
its rows in the same epilogue.
Under autograd (recommendflow_b200/training.py: batch statistics, dropout) a stage is ONE autograd node (`_TrainStage`):
the batch statistics are one var_mean pass, the normalisation is folded into the weights exactly as at inference (it is
an affine map per column for the batch too), and the forward GEMM and both backward GEMMs (dX = dZ W^T, dW = X^T dZ) run
on the same tcgen05 kernel; the BatchNormalization backward is applied to dX in closed form.  Shapes the tensor-core kernel
does not take (widths not multiples of 4, gelu) run as library GEMMs + elementwise torch ops.
All learned state is registered (Parameters / buffers), so it is part of `state_dict()`.
"""
import torch

from ... import dense_ops
from ..layers.attention_layers import Dense, library_matmul
from ..layers.preprocess_layers import Layer


def _selu(x):
    return torch.nn.functional.selu(x)


_ACT = {None: lambda x: x, "linear": lambda x: x, "relu": torch.relu, "selu": _selu, "tanh": torch.tanh,
        "sigmoid": torch.sigmoid, "gelu": torch.nn.functional.gelu}


class BatchNormalization(Layer):
    """Keras BatchNormalization: gamma * (x - mean) / sqrt(var + eps) + beta, moving statistics at inference,
    batch statistics (and a moving-average update) when `batch_stats` is on (Keras training=True)."""

    def __init__(self, epsilon=1e-3, momentum=0.99, name=None):
        super().__init__(name=name)
        self.epsilon = epsilon
        self.momentum = momentum
        self.batch_stats = False  # True during a training step: normalise with the batch's own statistics

    # ---- per-width state: gamma_<d>, beta_<d> (Parameters), moving_mean_<d>, moving_var_<d> (buffers) --------------
    def widths(self):
        return sorted(int(k.split("_")[1]) for k in self._parameters if k.startswith("gamma_"))

    def ensure(self, d, device=None):
        if f"gamma_{d}" not in self._parameters:
            self.register_parameter(f"gamma_{d}", torch.nn.Parameter(torch.ones(d, device=device), requires_grad=False))
            self.register_parameter(f"beta_{d}", torch.nn.Parameter(torch.zeros(d, device=device), requires_grad=False))
            self.register_buffer(f"moving_mean_{d}", torch.zeros(d, device=device))
            self.register_buffer(f"moving_var_{d}", torch.ones(d, device=device))
        return self

    def state(self, d):
        return (getattr(self, f"gamma_{d}"), getattr(self, f"beta_{d}"), getattr(self, f"moving_mean_{d}"),
                getattr(self, f"moving_var_{d}"))

    def set_weights(self, weights):
        gamma, beta, mean, var = (torch.as_tensor(w, dtype=torch.float32) for w in weights)
        d = int(gamma.numel())
        dev = self.state(d)[0].device if f"gamma_{d}" in self._parameters else (
            torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None)
        self.ensure(d, dev)
        with torch.no_grad():
            for dst, src in zip(self.state(d), (gamma, beta, mean, var)):
                dst.copy_(src.to(dst.device))

    def affine(self, d):
        """(scale, shift) of the inference-mode map x -> x * scale + shift for inputs of width d."""
        gamma, beta, mean, var = self.state(d)
        scale = gamma * torch.rsqrt(var + self.epsilon)
        return scale, beta - mean * scale

    def versions(self, d):
        return tuple((t.data_ptr(), t._version) for t in self.state(d))

    def call(self, x):
        d = x.shape[-1]
        self.ensure(d, x.device)
        gamma, beta, mean, var = self.state(d)
        if self.batch_stats:     # Keras training=True: batch mean / biased variance, moving statistics updated
            bmean, bvar = x.mean(dim=0), x.var(dim=0, unbiased=False)
            with torch.no_grad():
                mean.mul_(self.momentum).add_(bmean.detach() * (1 - self.momentum))
                var.mul_(self.momentum).add_(bvar.detach() * (1 - self.momentum))
            return (x - bmean) * (gamma * torch.rsqrt(bvar + self.epsilon)) + beta
        return (x - mean) * (gamma * torch.rsqrt(var + self.epsilon)) + beta


class _Activated(Layer):
    def __init__(self, units, activation):
        super().__init__(name="dense")
        if activation not in _ACT:
            raise ValueError(f"Unknown activation function: {activation}")
        self.dense = Dense(units)
        self.activation = activation

    def call(self, x):
        self.dense.build(x.shape[-1], x.device)
        if (dense_ops.DEFAULT_PRECISION == "tf32" and not self.dense.recording_grad(x)
                and dense_ops.dense_tc_ok(x, self.dense.kernel.shape[0], self.dense.units)):
            return dense_ops.dense_forward(x, self.dense.kernel_t(), self.dense.bias, self.activation)
        return _ACT[self.activation](library_matmul(x, self.dense.kernel) + self.dense.bias)


_SELU_SCALE, _SELU_ALPHA = 1.0507009873554805, 1.6732632423543772


class _TrainStage(torch.autograd.Function):
    """y = act(BN_batch(x) W + b) as one node.  BN_batch(x) = x * s + t with s = gamma * rsqrt(var + eps), t = beta - mean * s,
    so z = x (diag(s) W) + (t W + b): one tcgen05 GEMM on x itself.  Backward, with dZ = dY * act'(y) and db = colsum(dZ):
        dXhat = dZ W^T                                   (tcgen05; W [in, units] is already the [out, in] operand)
        dW    = Xhat^T dZ = diag(s) (X^T dZ) + t (x) db  (tcgen05 on the transposes; Xhat is never materialised)
        dbeta = colsum(dXhat), dgamma = colsum(dXhat * xn), dX = s / B * (B dXhat - dbeta - xn * dgamma),  xn = (x - mean) * rstd
    (Keras BatchNormalization, training=True: biased batch variance.)"""

    @staticmethod
    def forward(ctx, x, gamma, beta, kernel, bias, activation, eps, stats):
        x = x if x.stride(1) == 1 and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0 else x.contiguous()
        if gamma is not None:
            mean, var, xt = dense_ops.column_stats(x, want_transpose=True)        # one read of x: statistics + x^T for dW
            rstd = torch.rsqrt(var + eps)
            s = gamma * rstd
            t = beta - mean * s
            wt = (kernel.t() * s[None, :]).contiguous()
            b = torch.addmv(bias, kernel.t(), t)
            stats["mean"], stats["var"] = mean, var
        else:
            mean = rstd = s = t = None
            xt = x.t().contiguous()
            wt, b = kernel.t().contiguous(), bias
        y = dense_ops.dense_forward(x, wt, b, activation)
        ctx.activation, ctx.has_norm = activation, gamma is not None
        ctx.save_for_backward(x, xt, y, kernel, mean, rstd, s, t)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, xt, y, kernel, mean, rstd, s, t = ctx.saved_tensors
        dz, dzt, db = dense_ops.activation_backward(dy, y, ctx.activation)           # one read of dY, y
        dxh = dense_ops.dense_forward(dz, kernel, None, None)                        # [B, in]
        xtdz = dense_ops.dense_forward(xt, dzt, None, None)                          # [in, units]
        if not ctx.has_norm:
            return dxh, None, None, xtdz, db, None, None, None
        dw = xtdz.mul_(s[:, None])
        dw.addr_(t, db)
        dx, dgamma, dbeta = dense_ops.batchnorm_backward(dxh, x, mean, rstd, s)      # two reads of dXhat, x
        return dx, dgamma, dbeta, dw, db, None, None, None


class Dropout(Layer):
    """Keras Dropout(rate): identity unless `active` (a training step), then inverted dropout."""

    def __init__(self, rate, name=None):
        super().__init__(name=name)
        self.rate = rate
        self.active = False

    def call(self, x):
        return torch.nn.functional.dropout(x, self.rate, training=True) if self.active and self.rate > 0 else x


class Sequential(Layer):
    def __init__(self, layers, name=None):
        super().__init__(name=name)
        self.layers = list(layers)
        for i, l in enumerate(self.layers):
            if isinstance(l, torch.nn.Module):
                self.add_module(f"l{i}", l)
        self._folded = {}

    # ---- inference: [BatchNormalization, Dense + activation, Dropout] = one tensor-core launch ------------------------
    def _stages(self):
        """Group the layer list into (norm or None, _Activated) stages; None if the list has another shape."""
        stages, norm = [], None
        for l in self.layers:
            if isinstance(l, BatchNormalization) and norm is None:
                norm = l
            elif isinstance(l, _Activated):
                stages.append((norm, l))
                norm = None
            elif isinstance(l, Dropout):
                continue
            else:
                return None
        return stages if norm is None else None

    def _folded_weights(self, idx, norm, act, in_dim):
        """(weight_t [units, in], bias [units]) of stage idx with the inference-mode normalisation folded in:
        (x * s + t) W + b = x (diag(s) W) + (t W + b).  Cached until the kernel, bias or statistics change."""
        dense = act.dense
        key = (dense.kernel.data_ptr(), dense.kernel._version, dense.bias.data_ptr(), dense.bias._version,
               norm.versions(in_dim) if norm is not None else None)
        hit = self._folded.get(idx)
        if hit is None or hit[0] != key:
            with torch.no_grad():
                if norm is None:
                    wt, b = dense.kernel_t(), dense.bias.detach()
                else:
                    scale, shift = norm.affine(in_dim)
                    wt = (dense.kernel.detach().t() * scale[None, :]).contiguous()
                    b = dense.bias.detach() + shift @ dense.kernel.detach()
            hit = (key, wt, b)
            self._folded[idx] = hit
        return hit[1], hit[2]

    def _fusable(self, x, stages):
        if stages is None or dense_ops.DEFAULT_PRECISION != "tf32" or not x.is_cuda or x.dtype != torch.float32 or x.dim() != 2:
            return False
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            return False
        for l in self.layers:
            if (isinstance(l, BatchNormalization) and l.batch_stats) or (isinstance(l, Dropout) and l.active and l.rate > 0):
                return False
        d = x.shape[-1]
        for norm, act in stages:
            if d % 4 or act.dense.units % 4:
                return False
            d = act.dense.units
        return True

    def _trainable_on_tc(self, x, stages):
        if stages is None or dense_ops.DEFAULT_PRECISION != "tf32" or not x.is_cuda or x.dtype != torch.float32 or x.dim() != 2:
            return False
        if not torch.is_grad_enabled() or x.shape[0] % 4:
            return False
        d = x.shape[-1]
        for norm, act in stages:
            if d % 4 or act.dense.units % 4 or act.activation == "gelu":
                return False
            d = act.dense.units
        return True

    def _train_call(self, x, stages):
        """The gradient-recording path: one `_TrainStage` node per [norm, Dense + activation], then the stage's Dropout."""
        drops = [l for l in self.layers if isinstance(l, Dropout)]
        for idx, (norm, act) in enumerate(stages):
            d = x.shape[-1]
            act.dense.build(d, x.device)
            gamma = beta = None
            stats = {}
            eps = 0.0
            if norm is not None:
                norm.ensure(d, x.device)
                gamma, beta, mmean, mvar = norm.state(d)
                eps = norm.epsilon
                if not norm.batch_stats:      # moving statistics under autograd (fine-tuning with frozen statistics): plain layers
                    x = act(norm(x))
                    x = drops[idx](x) if idx < len(drops) else x
                    continue
            x = _TrainStage.apply(x, gamma, beta, act.dense.kernel, act.dense.bias, act.activation, eps, stats)
            if norm is not None:
                with torch.no_grad():
                    mmean.mul_(norm.momentum).add_(stats["mean"] * (1 - norm.momentum))
                    mvar.mul_(norm.momentum).add_(stats["var"] * (1 - norm.momentum))
            if idx < len(drops):
                x = drops[idx](x)
        return x

    def call(self, x, l2_normalize=False):
        """l2_normalize: divide every output row by max(||row||, 1e-12) -- fused into the last stage's epilogue on
        the tensor-core path (the towers' embedding_norm)."""
        stages = self._stages()
        if not self._fusable(x, stages) and self._trainable_on_tc(x, stages):
            x = self._train_call(x, stages)
            return torch.nn.functional.normalize(x, dim=1, eps=1e-12) if l2_normalize else x
        if self._fusable(x, stages) and (not l2_normalize or stages[-1][1].dense.units <= 256):
            for idx, (norm, act) in enumerate(stages):
                d = x.shape[-1]
                act.dense.build(d, x.device)
                if norm is not None:
                    norm.ensure(d, x.device)
                wt, b = self._folded_weights(idx, norm, act, d)
                x = dense_ops.dense_forward(x, wt, b, act.activation, l2_normalize and idx == len(stages) - 1)
            return x
        for layer in self.layers:
            x = layer(x)
        return torch.nn.functional.normalize(x, dim=1, eps=1e-12) if l2_normalize else x


def create_mlp(hidden_units, dropout_rate, activation, normalization_layer, name=None):
    layers = []
    for units in hidden_units:
        if normalization_layer is not None:
            layers.append(normalization_layer)
        layers.append(_Activated(units, activation))
        layers.append(Dropout(dropout_rate))
    return Sequential(layers, name=name)
