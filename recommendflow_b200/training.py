"""One optimisation step of the two-tower recall model, the way the reference's `model.fit` drives it
(/root/reference/example/recall_search/train.py:97-104: `tf.keras.optimizers.Adam(learning_rate)` on
every variable, loss added by the model).

What runs where:
  * forward: ONE fused hash + gather + pool launch for all features (ids kept), the SDPA encoder and the
    in-batch softmax loss on the CUDA kernels, tower GEMMs on cuBLAS;
  * backward: rf_inbatch_softmax_ce_backward and rf_sdpa_backward through torch.autograd Functions, the
    tower / projection GEMMs through torch's own autograd (library code);
  * update: dense variables by `KerasAdam` below -- tf.keras.optimizers.Adam's exact update
    w -= lr_t * m / (sqrt(v) + eps), lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t), the same form rf_bag_adam.cu applies to
    the tables (torch.optim.Adam adds eps to sqrt(v_hat), a ~30x different effective eps at step 1) -- every
    embedding table by rf_bag_backward_adam (Keras' sparse Adam: duplicates summed, all rows decay; `lazy=True`
    restricts the update to the gathered rows).
BatchNormalization uses batch statistics and Dropout is active during the step, as with Keras training=True.
Single GPU (replicated tables); the reverse exchange of the row-sharded path is not built yet.
"""
import torch

from .backend.blocks.mlp import BatchNormalization, Dropout
from .backend.layers.preprocess_layers import DoubleHashingEmbedding
from .bag_ops import BagAdamGroup


class KerasAdam(object):
    """tf.keras.optimizers.Adam on dense variables (foreach tensor ops; O(#parameters) glue, not a hot path).
    m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  w -= lr * sqrt(1 - b2^t) / (1 - b1^t) * m / (sqrt(v) + eps)."""

    def __init__(self, params, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.params = [p for p in params]
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon = learning_rate, beta_1, beta_2, epsilon
        self.iterations = 0
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        # device-resident step (`capturable`): t and lr_t live in CUDA memory and are advanced by tensor ops, so that a step
        # recorded into a CUDA graph keeps counting when it is replayed
        self.capturable = False
        self.step_t = self.lr_t = None

    def make_capturable(self):
        dev = self.params[0].device
        self.capturable = True
        self.step_t = torch.full((), float(self.iterations), dtype=torch.float64, device=dev)
        self.lr_t = torch.zeros((), dtype=torch.float32, device=dev)
        return self

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    @torch.no_grad()
    def step(self):
        self.iterations += 1
        t = self.iterations
        if self.capturable:
            self.step_t += 1.0
            lr = self.learning_rate * torch.sqrt(1.0 - torch.pow(self.beta_2, self.step_t)) / (1.0 - torch.pow(self.beta_1, self.step_t))
            self.lr_t.copy_(lr)
        else:
            lr_t = self.learning_rate * (1.0 - self.beta_2 ** t) ** 0.5 / (1.0 - self.beta_1 ** t)
        idx = [i for i, p in enumerate(self.params) if p.grad is not None]
        if not idx:
            return
        ps, gs = [self.params[i] for i in idx], [self.params[i].grad for i in idx]
        ms, vs = [self.m[i] for i in idx], [self.v[i] for i in idx]
        torch._foreach_mul_(ms, self.beta_1)
        torch._foreach_add_(ms, gs, alpha=1.0 - self.beta_1)
        torch._foreach_mul_(vs, self.beta_2)
        torch._foreach_addcmul_(vs, gs, gs, value=1.0 - self.beta_2)
        den = torch._foreach_sqrt(vs)
        torch._foreach_add_(den, self.epsilon)
        if self.capturable:
            upd = torch._foreach_div(ms, den)
            torch._foreach_mul_(upd, self.lr_t)
            torch._foreach_sub_(ps, upd)
        else:
            torch._foreach_addcdiv_(ps, ms, den, value=-lr_t)

    def state_dict(self):
        return {"iterations": self.iterations, "m": [t.clone() for t in self.m], "v": [t.clone() for t in self.v]}

    def load_state_dict(self, state):
        self.iterations = int(state["iterations"])
        for dst, src in zip(self.m, state["m"]):
            dst.copy_(src)
        for dst, src in zip(self.v, state["v"]):
            dst.copy_(src)


class RecallSdpaTrainer(object):
    def __init__(self, model, learning_rate=1e-4, lazy_embedding_adam=False):
        self.model = model
        self.learning_rate = learning_rate
        self.lazy = lazy_embedding_adam
        self.dense_opt = None
        self.bag_opts = {}
        self._fused_args = {}
        self.iterations = 0

    # ---- helpers ---------------------------------------------------------------------------------
    def _modules(self, kind):
        # the module tree (hundreds of embedding layers) is walked once per kind, not four times per step
        cache = self.__dict__.setdefault("_modules_cache", {})
        key = (kind, len(self.model.preprocessor))
        if key not in cache:
            cache[key] = [m for m in self.model.modules() if isinstance(m, kind)]
        return cache[key]

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
        beta).  The embedding tables are registered parameters too (state_dict), but they are updated by the fused
        sparse Adam kernels, never by autograd."""
        tables = set()
        for layer in self.model.preprocessor.values():
            tables.update(id(p) for p in layer.parameters())
        params, seen = [], set()
        for p in self.model.parameters():
            if id(p) in tables or id(p) in seen:
                continue
            seen.add(id(p))
            p.requires_grad_(True)
            params.append(p)
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

    def _fused_layout(self, names):
        """Column layout of the fused bag output for these features (cached: 228 features make it ~15 ms of Python otherwise)."""
        key = (tuple(names), len(self.model.preprocessor))
        hit = self.__dict__.get("_layout_cache")
        if hit is None or hit[0] != key:
            fusable = set(self.model.preprocessor.fused_names())
            hit = (key, self.model.preprocessor.output_layout([n for n in names if n in fusable]))
            self.__dict__["_layout_cache"] = hit
        return hit[1]

    def _forward(self, batch, y_true, behaviour, ids):
        names = self.model.user_cols + self.model.ad_cols
        embs = self.model.preprocessor.forward_all(batch, names=names, keep_ids=ids)
        missing = [n for n in names if n not in ids]
        if missing:
            raise NotImplementedError(f"features {missing} do not pool (combiner null / first / last): no training path")
        leaf = embs["__fused__"].detach().requires_grad_(True)
        layout, _ = self._fused_layout(names)
        # the fused buffer holds the user features first, then the ad features (names order): each tower's input is ONE
        # column window of the leaf -- two slices on the autograd tape instead of one per feature
        ucols = sum(layout[n][1] for n in self.model.user_cols)
        u, a = self.model.towers_from_fused(leaf[:, :ucols], leaf[:, ucols:], behaviour)
        return self.model.loss_fun(y_true, u, a), leaf, layout

    # ---- the step --------------------------------------------------------------------------------
    def train_step(self, batch, y_true, behaviour=None):
        """batch: {feature: StringColumn | tensor | lists}; returns the loss (0-dim CUDA tensor) before the update."""
        if self.dense_opt is None:                      # variables are created lazily by the first forward
            with torch.no_grad():
                self._forward(batch, y_true, behaviour, {})
            self.dense_opt = KerasAdam(self._dense_variables(), learning_rate=self.learning_rate)
        self._set_training(True)
        try:
            ids = {}
            loss, leaf, layout = self._forward(batch, y_true, behaviour, ids)
            self.dense_opt.zero_grad()
            loss.backward()
        finally:
            self._set_training(False)
        self.dense_opt.step()
        grad = leaf.grad
        for dim, (group, members) in self._bag_groups(layout).items():
            # the forward's cached plan keeps the id buffers in place: the per-table views are rebuilt only when they moved
            key = tuple((ids[name][0].data_ptr(), ids[name][1]) for name, t, _ in members if t == 0)
            hit = self._fused_args.get(dim)
            if hit is None or hit[0] != key:
                id_list, cols, combs, lens = [], [], [], []
                for name, t, bag in members:
                    layer = self.model.preprocessor[name]
                    combiner = layer.combiner if isinstance(layer, DoubleHashingEmbedding) else layer.pooling
                    if combiner not in ("sum", "avg"):
                        raise NotImplementedError(f"feature {name}: backward is implemented for sum / avg pooling, not {combiner}")
                    rows, bag_len = ids[name]
                    id_list.append(rows[t])
                    cols.append(layout[name][0] + t * dim)
                    combs.append(combiner)
                    lens.append(bag_len)
                hit = (key, id_list, cols, combs, lens)
                self._fused_args[dim] = hit
            group.apply_fused(hit[1], grad, hit[2], hit[3], hit[4], grad.shape[0],
                              lr_t=self.dense_opt.lr_t if self.dense_opt.capturable else None)
        self.iterations += 1
        return loss.detach()

    # ---- checkpointing: model.state_dict() holds the variables; this holds the optimizer slots -------------------
    def state_dict(self):
        """Adam moments of every dense variable and every embedding table + the iteration counters (what
        tf.keras' ModelCheckpoint(save_weights_only=False) keeps beside the weights)."""
        return {"iterations": self.iterations,
                "dense": None if self.dense_opt is None else self.dense_opt.state_dict(),
                "bags": {dim: group.state_dict() for dim, (group, _) in self.bag_opts.items()}}

    def load_state_dict(self, state, example_batch=None):
        """Restore after the optimizers exist (they are created lazily by the first step: pass `example_batch` =
        (batch, y_true, behaviour) to build them without taking a step)."""
        if self.dense_opt is None:
            if example_batch is None:
                raise RuntimeError("the optimizers are built by the first step; pass example_batch to build them now")
            batch, y_true, behaviour = example_batch
            with torch.no_grad():
                self._forward(batch, y_true, behaviour, {})
            self.dense_opt = KerasAdam(self._dense_variables(), learning_rate=self.learning_rate)
        if not self.bag_opts and state["bags"]:
            names = self.model.user_cols + self.model.ad_cols
            layout, _ = self._fused_layout(names)
            self._bag_groups(layout)
        self.iterations = int(state["iterations"])
        if state["dense"] is not None:
            self.dense_opt.load_state_dict(state["dense"])
        for dim, st in state["bags"].items():
            self.bag_opts[dim][0].load_state_dict(st)


class GraphedTrainStep(object):
    """One training step recorded into a CUDA graph and replayed: the eager step issues ~350 launches from Python (forward
    plan, autograd, two optimizers) and is bound by the host; the replay costs one launch.

    The tensors of `batch`, `y_true` and `behaviour` given here are the graph's static inputs: refill them IN PLACE
    (`.copy_`) between replays -- the standard CUDA-graph contract.  Shapes are fixed.  The step counters of both optimizers
    live on the device (`KerasAdam.make_capturable`), so the bias correction keeps advancing under replay.
        step = GraphedTrainStep(trainer, batch, y, behaviour);  loss = step()      # loss: 0-dim CUDA tensor, overwritten per replay
    """

    def __init__(self, trainer, batch, y_true, behaviour=None, warmup=3):
        self.trainer = trainer
        if trainer.dense_opt is None:
            trainer.train_step(batch, y_true, behaviour)
        if not trainer.dense_opt.capturable:
            trainer.dense_opt.make_capturable()
        dev = trainer.dense_opt.params[0].device
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):          # plans, workspaces and descriptor pools are built outside the capture
                trainer.train_step(batch, y_true, behaviour)
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = trainer.train_step(batch, y_true, behaviour)
        # recording ran the Python of one step (host counters moved) but none of its kernels: take that step back
        self._bump(-1)
        self._keep = (batch, y_true, behaviour)

    def _bump(self, by):
        t = self.trainer
        t.iterations += by
        t.dense_opt.iterations += by
        for group, _ in t.bag_opts.values():
            group.iterations += by

    def __call__(self):
        self.graph.replay()
        self._bump(1)
        return self.loss
