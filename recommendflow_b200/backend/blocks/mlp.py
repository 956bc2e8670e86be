"""Tower MLP factory with the reference's signature (/root/reference/backend/blocks/mlp.py:4-15):
`create_mlp(hidden_units, dropout_rate, activation, normalization_layer)` -> [norm, Dense, Dropout] * n.

Reference quirk kept: the SAME normalization layer instance is appended before every Dense (:11), so
one BatchNormalization would be applied to inputs of different widths -- Keras builds it on the first
width and fails on the second unless all widths are equal.  Here the shared instance keeps one set of
statistics per distinct width (the only way the reference's towers [1024, 512, 256] can run at all).
Dense layers are library GEMMs (cuBLAS via torch).  Inference mode (moving statistics, no dropout) unless a
training step switches `BatchNormalization.batch_stats` / `Dropout.active` on (recommendflow_b200/training.py).
"""
import torch

from ..layers.attention_layers import Dense
from ..layers.preprocess_layers import Layer


def _selu(x):
    return torch.nn.functional.selu(x)


_ACT = {None: lambda x: x, "linear": lambda x: x, "relu": torch.relu, "selu": _selu, "tanh": torch.tanh,
        "sigmoid": torch.sigmoid, "gelu": torch.nn.functional.gelu}


class BatchNormalization(Layer):
    """Keras BatchNormalization at inference: gamma * (x - moving_mean) / sqrt(moving_var + eps) + beta."""

    def __init__(self, epsilon=1e-3, momentum=0.99, name=None):
        super().__init__(name=name)
        self.epsilon = epsilon
        self.momentum = momentum
        self.stats = {}          # width -> (gamma, beta, moving_mean, moving_var)
        self.batch_stats = False  # True during a training step: normalise with the batch's own statistics

    def set_weights(self, weights):
        gamma, beta, mean, var = (torch.as_tensor(w, dtype=torch.float32) for w in weights)
        self.stats[int(gamma.numel())] = (gamma, beta, mean, var)

    def call(self, x):
        d = x.shape[-1]
        if d not in self.stats:
            self.stats[d] = (torch.ones(d), torch.zeros(d), torch.zeros(d), torch.ones(d))
        gamma, beta, mean, var = (t.to(x.device) for t in self.stats[d])
        if self.batch_stats:     # Keras training=True: batch mean / biased variance, moving statistics updated
            bmean, bvar = x.mean(dim=0), x.var(dim=0, unbiased=False)
            with torch.no_grad():
                mean = mean * self.momentum + bmean.detach() * (1 - self.momentum)
                var = var * self.momentum + bvar.detach() * (1 - self.momentum)
            self.stats[d] = (gamma, beta, mean, var)
            return (x - bmean) * (gamma * torch.rsqrt(bvar + self.epsilon)) + beta
        self.stats[d] = (gamma, beta, mean, var)
        return (x - mean) * (gamma * torch.rsqrt(var + self.epsilon)) + beta

    def trainable(self):
        """gamma / beta of every width seen so far, as leaves that require grad."""
        out = []
        for d, (gamma, beta, mean, var) in list(self.stats.items()):
            gamma, beta = gamma.detach().requires_grad_(True), beta.detach().requires_grad_(True)
            self.stats[d] = (gamma, beta, mean, var)
            out += [gamma, beta]
        return out


class Sequential(Layer):
    def __init__(self, layers, name=None):
        super().__init__(name=name)
        self.layers = list(layers)
        for i, l in enumerate(self.layers):
            if isinstance(l, torch.nn.Module):
                self.add_module(f"l{i}", l)

    def call(self, x):
        for layer in self.layers:
            x = layer(x)
        return x


class _Activated(Layer):
    def __init__(self, units, activation):
        super().__init__(name="dense")
        if activation not in _ACT:
            raise ValueError(f"Unknown activation function: {activation}")
        self.dense = Dense(units)
        self.activation = activation

    def call(self, x):
        return _ACT[self.activation](self.dense(x))


class Dropout(Layer):
    """Keras Dropout(rate): identity unless `active` (a training step), then inverted dropout."""

    def __init__(self, rate, name=None):
        super().__init__(name=name)
        self.rate = rate
        self.active = False

    def call(self, x):
        return torch.nn.functional.dropout(x, self.rate, training=True) if self.active and self.rate > 0 else x


def create_mlp(hidden_units, dropout_rate, activation, normalization_layer, name=None):
    layers = []
    for units in hidden_units:
        layers.append(normalization_layer)
        layers.append(_Activated(units, activation))
        layers.append(Dropout(dropout_rate))
    return Sequential(layers, name=name)
