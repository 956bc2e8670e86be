"""One optimisation step of the two-tower recall model, the way the reference's `model.fit` drives it
(/root/reference/example/recall_search/train.py:97-104: `tf.keras.optimizers.Adam(learning_rate)` on
every variable, loss added by the model).

What runs where:
  * forward: ONE fused hash + gather + pool launch for all features (ids kept), the SDPA encoder and the
    in-batch softmax loss on the CUDA kernels, tower GEMMs on cuBLAS;
  * backward: rf_inbatch_softmax_ce_backward and rf_sdpa_backward through torch.autograd Functions, the
    tower / projection GEMMs through torch's own autograd (library code);
  * update: dense variables by torch.optim.Adam with Keras' hyper-parameters (eps = 1e-7), every embedding
    table by rf_bag_backward_adam (Keras' sparse Adam: duplicates summed, all rows decay; `lazy=True`
    restricts the update to the gathered rows).
BatchNormalization uses batch statistics and Dropout is active during the step, as with Keras training=True.
Single GPU (replicated tables); the reverse exchange of the row-sharded path is not built yet.
"""
import torch

from .backend.blocks.mlp import BatchNormalization, Dropout
from .backend.layers.preprocess_layers import DoubleHashingEmbedding
from .bag_ops import BagAdamGroup


class RecallSdpaTrainer(object):
    def __init__(self, model, learning_rate=1e-4, lazy_embedding_adam=False):
        self.model = model
        self.learning_rate = learning_rate
        self.lazy = lazy_embedding_adam
        self.dense_opt = None
        self.bag_opts = {}
        self.iterations = 0

    # ---- helpers ---------------------------------------------------------------------------------
    def _modules(self, kind):
        return [m for m in self.model.modules() if isinstance(m, kind)]

    def _set_training(self, flag):
        seen = set()
        for m in self._modules(BatchNormalization):
            if id(m) not in seen:
                m.batch_stats = flag
                seen.add(id(m))
        for m in self._modules(Dropout):
            m.active = flag

    def _dense_variables(self):
        """Every dense variable of the model (tower / projection kernels and biases, BatchNormalization gamma and
        beta).  The embedding tables live in the preprocessing layers, outside `model.parameters()`."""
        params = list(self.model.parameters())
        for p in params:
            p.requires_grad_(True)
        seen = set()
        for bn in self._modules(BatchNormalization):
            if id(bn) not in seen:
                params += bn.trainable()
                seen.add(id(bn))
        return params

    def _bag_groups(self, layout):
        """One BagAdamGroup per embedding width: every table of that width is updated in ONE pass per step."""
        if not self.bag_opts:
            by_dim = {}
            for name, (col, width) in layout.items():
                layer = self.model.preprocessor[name]
                bags = (layer.emb1, layer.emb2) if isinstance(layer, DoubleHashingEmbedding) else (layer.embedding,)
                for t, bag in enumerate(bags):
                    by_dim.setdefault(bag.output_dim, []).append((name, t, bag))
            for dim, members in by_dim.items():
                group = BagAdamGroup([bag.embeddings.data for _, _, bag in members], learning_rate=self.learning_rate,
                                     lazy=self.lazy)
                self.bag_opts[dim] = (group, members)
        return self.bag_opts

    def _forward(self, batch, y_true, behaviour, ids):
        names = self.model.user_cols + self.model.ad_cols
        embs = self.model.preprocessor.forward_all(batch, names=names, keep_ids=ids)
        missing = [n for n in names if n not in ids]
        if missing:
            raise NotImplementedError(f"features {missing} do not pool (combiner null / first / last): no training path")
        leaf = embs["__fused__"].detach().requires_grad_(True)
        layout, _ = self.model.preprocessor.output_layout([n for n in names if n in set(self.model.preprocessor.fused_names())])
        views = {n: leaf[:, col:col + width] for n, (col, width) in layout.items()}
        u, a = self.model.towers_from_embeddings(views, behaviour)
        return self.model.loss_fun(y_true, u, a), leaf, layout

    # ---- the step --------------------------------------------------------------------------------
    def train_step(self, batch, y_true, behaviour=None):
        """batch: {feature: StringColumn | tensor | lists}; returns the loss (0-dim CUDA tensor) before the update."""
        if self.dense_opt is None:                      # variables are created lazily by the first forward
            with torch.no_grad():
                self._forward(batch, y_true, behaviour, {})
            self.dense_opt = torch.optim.Adam(self._dense_variables(), lr=self.learning_rate, betas=(0.9, 0.999), eps=1e-7)
        self._set_training(True)
        try:
            ids = {}
            loss, leaf, layout = self._forward(batch, y_true, behaviour, ids)
            self.dense_opt.zero_grad(set_to_none=True)
            loss.backward()
        finally:
            self._set_training(False)
        self.dense_opt.step()
        grad = leaf.grad
        for dim, (group, members) in self._bag_groups(layout).items():
            updates = []
            for name, t, bag in members:
                layer = self.model.preprocessor[name]
                combiner = layer.combiner if isinstance(layer, DoubleHashingEmbedding) else layer.pooling
                if combiner not in ("sum", "avg"):
                    raise NotImplementedError(f"feature {name}: backward is implemented for sum / avg pooling, not {combiner}")
                rows, bag_len = ids[name]
                col = layout[name][0] + t * dim
                updates.append((rows[t], grad[:, col:col + dim], combiner, bag_len, None))
            group.apply(updates, grad.shape[0])
        self.iterations += 1
        return loss.detach()
