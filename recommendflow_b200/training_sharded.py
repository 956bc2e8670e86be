"""C5 (BASELINE.json configs[4]): one optimisation step of the two-tower recall model on G GPUs with the embedding
tables ROW-SHARDED across them and the dense towers DATA-PARALLEL (either transport of the sharded bags: nccl or p2p).

What the reference does instead (/root/reference/example/ranking_search/train.py:93-104 and
backend/utils/gpu_utils.py:13-14): `tf.distribute.MirroredStrategy` -- every table replicated on every GPU, one
`tf.keras.optimizers.Adam` on all variables, gradients all-reduced.  Replication stops working once the tables no
longer fit one GPU (1 B rows), which is what north_star's config 5 asks for.  Here, per step and rank:

  forward   every sharded feature: route ids -> owners pool -> partials back -> combine  (sharded.ShardedEmbeddingBag)
            towers on the rank's own B samples (replicated dense variables)
            in-batch softmax over ALL G*B docs: the doc embeddings are all-gathered and the [B, G*B] logits are taken one
            [B, B] block at a time by the same kernels the single-GPU loss uses (the row log-sum-exp is merged over
            the blocks; positives sit on the diagonal of the rank's own block); loss = mean over the global batch
  backward  dQ = sum over blocks of C_g A_g,  dA_g = C_g^T Q per block, reduce-scattered (summed) to the docs' owners;
            torch autograd through the towers; dense gradients all-reduced (mean); pooled-bag gradients go back through
            `ShardedEmbeddingBag.apply_adam` (reverse exchange + Keras Adam on the owners' rows)
  update    `training.KerasAdam` on the dense variables (identical on every rank), the fused sparse Adam on the shards.

The collectives are torch.distributed (NCCL on GPUs; gloo in the CPU tests).  Compute ops are injected (`ops`) so that
the choreography can be tested under gloo with CPU stand-ins (tests/shard_util.py); the default ops are the CUDA kernels.
"""
import torch
import torch.distributed as dist

from . import dense_ops
from .backend.blocks.mlp import BatchNormalization, Dropout
from .training import KerasAdam


class CudaLossOps(object):
    """The [B, B] block primitives of the all-gathered in-batch softmax on the CUDA kernels."""

    def block_lse(self, q, a, scale):
        """log sum_j exp(scale * q_i . a_j) for every row i (rf_inbatch_rowstats[_tc])."""
        return dense_ops.inbatch_rowstats(q, a, scale=scale, want=("lse",))["lse"]

    def rowdot(self, q, a):
        return dense_ops.inbatch_rowstats(q, a, want=("diag",), precision="fp32")["diag"]

    def block_grads(self, q, a, y, lse, scale, upstream, own_block):
        """(dQ, dA) of one block given the FULL-row lse (rf_inbatch_softmax_ce_backward_block)."""
        return dense_ops.inbatch_softmax_ce_backward(q, a, y, lse, scale, upstream, True, True, positives_on_diagonal=own_block)


class AllGatherInbatchCE(torch.autograd.Function):
    """batch_neg_sample_scaled_multi_class_ce_loss (match_losses.py:150-165) with the negatives of the WHOLE global batch:
    returns this rank's mean over its B rows; the mean of that over ranks is the global-batch loss."""

    @staticmethod
    def forward(ctx, y, q, a, scale, group, ops):
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        q, a = q.contiguous(), a.contiguous()
        blocks = [torch.empty_like(a) for _ in range(world)]
        dist.all_gather(blocks, a, group=group)
        lse_blocks = torch.stack([ops.block_lse(q, blocks[g], scale) for g in range(world)])     # [G, B]
        lse = torch.logsumexp(lse_blocks, dim=0)
        diag = ops.rowdot(q, a)
        ctx.save_for_backward(y, q, lse, *blocks)
        ctx.scale, ctx.group, ctx.ops, ctx.rank, ctx.world = float(scale), group, ops, rank, world
        return torch.mean(-(scale * diag - lse) * y)

    @staticmethod
    def backward(ctx, grad_loss):
        y, q, lse, *blocks = ctx.saved_tensors
        dq = torch.zeros_like(q)
        da_all = []
        for g in range(ctx.world):              # upstream = 1 in the kernels, applied on the device below (no host sync)
            gq, ga = ctx.ops.block_grads(q, blocks[g], y, lse, ctx.scale, 1.0, g == ctx.rank)
            dq += gq
            da_all.append(ga.mul_(grad_loss))
        dq.mul_(grad_loss)
        # every rank holds a gradient for every rank's docs: sum them at the docs' owner
        if dist.get_backend(ctx.group) == "gloo":          # gloo has no reduce_scatter: all-reduce the stack, keep the own slice
            stack = torch.stack(da_all)
            dist.all_reduce(stack, op=dist.ReduceOp.SUM, group=ctx.group)
            da = stack[ctx.rank].contiguous()
        else:
            da = torch.empty_like(blocks[ctx.rank])
            dist.reduce_scatter(da, da_all, op=dist.ReduceOp.SUM, group=ctx.group)
        return None, dq, da, None, None, None


class ShardedRecallTrainer(object):
    """user_bags / ad_bags: {feature name: ShardedEmbeddingBag (nccl or p2p transport)}; user_tower / ad_tower: `create_mlp`
    Sequentials (replicated: construct them from the same seed on every rank).  batch: {feature name: keys}."""

    def __init__(self, user_bags, ad_bags, user_tower, ad_tower, learning_rate=1e-4, scale=20.0, group=None, loss_ops=None,
                 lazy_embedding_adam=False):
        self.user_bags, self.ad_bags = dict(user_bags), dict(ad_bags)
        self.user_tower, self.ad_tower = user_tower, ad_tower
        self.learning_rate, self.scale, self.lazy = learning_rate, scale, lazy_embedding_adam
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.loss_ops = loss_ops if loss_ops is not None else CudaLossOps()
        self.dense_opt = None
        self.iterations = 0

    def _dense_modules(self):
        return [m for tower in (self.user_tower, self.ad_tower) for m in tower.modules()]

    def _set_training(self, flag):
        for m in self._dense_modules():
            if isinstance(m, BatchNormalization):
                m.batch_stats = flag
            elif isinstance(m, Dropout):
                m.active = flag

    def _dense_variables(self):
        params, seen = [], set()
        for tower in (self.user_tower, self.ad_tower):
            for p in tower.parameters():
                if id(p) not in seen:
                    seen.add(id(p))
                    p.requires_grad_(True)
                    params.append(p)
        return params

    def _pooled(self, bags, batch):
        outs = [bag(batch[name]) for name, bag in bags.items()]
        x = torch.cat(outs, dim=1) if len(outs) > 1 else outs[0]
        return x.detach().requires_grad_(True)

    def forward(self, batch, y_true):
        """(local loss, pooled user input leaf, pooled ad input leaf)."""
        u_in, a_in = self._pooled(self.user_bags, batch), self._pooled(self.ad_bags, batch)
        u = torch.nn.functional.normalize(self.user_tower(u_in), dim=1, eps=1e-12)
        a = torch.nn.functional.normalize(self.ad_tower(a_in), dim=1, eps=1e-12)
        y = torch.as_tensor(y_true, dtype=torch.float32, device=u.device).reshape(-1)
        return AllGatherInbatchCE.apply(y, u, a, self.scale, self.group, self.loss_ops), u_in, a_in

    def train_step(self, batch, y_true):
        """One step; returns the GLOBAL-batch loss (mean over ranks) before the update."""
        if self.dense_opt is None:
            with torch.no_grad():          # builds the lazily created dense variables (same collectives on every rank)
                self.forward(batch, y_true)
            self.dense_opt = KerasAdam(self._dense_variables(), learning_rate=self.learning_rate)
        self._set_training(True)
        try:
            loss, u_in, a_in = self.forward(batch, y_true)
            self.dense_opt.zero_grad()
            loss.backward()
        finally:
            self._set_training(False)
        # d(global loss)/d theta = mean over ranks of the per-rank sums: all-reduce(SUM) / G, flattened into one buffer
        grads = [p.grad for p in self.dense_opt.params if p.grad is not None]
        if grads:
            flat = torch.cat([g.reshape(-1) for g in grads])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            flat /= self.world
            off = 0
            for g in grads:
                g.copy_(flat[off:off + g.numel()].view_as(g))
                off += g.numel()
        self.dense_opt.step()
        # sparse side: the gradient of the global loss w.r.t. this rank's pooled bags, back through the exchange
        for bags, leaf in ((self.user_bags, u_in), (self.ad_bags, a_in)):
            col = 0
            for name, bag in bags.items():
                g = leaf.grad[:, col:col + bag.output_dim] / self.world
                bag.apply_adam(g.contiguous(), learning_rate=self.learning_rate, lazy=self.lazy)
                col += bag.output_dim
        self.iterations += 1
        total = loss.detach().clone()
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=self.group)
        return total / self.world
