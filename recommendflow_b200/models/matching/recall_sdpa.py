"""Two-tower recall forward: hashed features -> fused bags -> (SDPA behaviour-sequence encoder) ->
tower MLPs -> l2-normalise -> in-batch softmax loss.

This is the "recall-SDPA" forward that conf/base_recall_sdpa.yaml describes.  The reference's own
model for that config, `models/matching/dssm.py:Dssm`, is an unfinished stub (its `call` never uses
its preprocessor or towers, :38-60), so this class composes the reference's building blocks the way
its working model does (`models/matching/que2search.py:68-79,114-140`: per-feature
`self.preprocessor[name](batch[name])`, concat per tower, tower MLP, `self.loss(y, q, a)`), with
Dssm's tower definition (`create_mlp([1024, 512, 256], 0.3, "selu", BatchNormalization(1e-6))`,
dssm.py:25-26) and an optional `MultiHeadAttention` encoder (attention_layers.py:137-168) over a
behaviour sequence, mean-pooled over the sequence, on the user side.

What runs where: every hashed feature of a batch goes through ONE fused kernel launch
(`forward_all`); SDPA and the B x B logits run on the tcgen05 kernels; tower GEMMs are cuBLAS.
"""
import torch

from ...backend.blocks.mlp import BatchNormalization, create_mlp
from ...backend.layers.attention_layers import MultiHeadAttention
from ...backend.lossess import match_losses
from ...backend.utils.preprocess_utils import get_preprocess_layers


class RecallSdpa(torch.nn.Module):
    def __init__(self, feature_conf, loss=None, tower_units=(1024, 512, 256), behaviour_dim=None, num_heads=1,
                 global_l2_norm=False, name="recall_sdpa"):
        super().__init__()
        self._name = name
        self.feature_conf = feature_conf
        self.preprocessor = get_preprocess_layers(feature_conf)
        # the mapping above keeps the reference's `layers[name](batch[name])` convention; registering the same layers
        # here makes every embedding table part of state_dict() / .to() (feature names may hold dots: keys are indexed)
        self.preprocess_layers = torch.nn.ModuleList(list(self.preprocessor.values()))
        self.user_cols = [f.name for f in feature_conf.features.get_features(tower="user") if f.name in self.preprocessor]
        self.ad_cols = [f.name for f in feature_conf.features.get_features(tower="ad") if f.name in self.preprocessor]
        self.user_dense = create_mlp(list(tower_units), 0.3, "selu", BatchNormalization(epsilon=1e-6), name="user_dense_tower")
        self.ad_dense = create_mlp(list(tower_units), 0.3, "selu", BatchNormalization(epsilon=1e-6), name="ad_dense_tower")
        self.seq_encoder = MultiHeadAttention(behaviour_dim, num_heads) if behaviour_dim else None
        self.loss_fun = loss or match_losses.batch_neg_sample_scaled_multi_class_ce_loss
        # Dssm.embedding_norm calls K.l2_normalize(x) with no axis = one norm over the whole batch
        # (dssm.py:36); per-row normalisation is what the in-batch softmax needs, so it is the default
        self.global_l2_norm = global_l2_norm
        # two-stream overlap of the sequence encoder with the fused bag launch: measured SLOWER on B200 (0.66 vs 0.61 ms per C3
        # forward as a graph, 0.94 vs 0.65 ms eager): the SDPA kernel is persistent with one 200 KB CTA per SM, so the bag CTAs cannot
        # co-reside and the two only interleave.  Kept as an option.
        self.overlap_encoder = False

    @property
    def name(self):
        return self._name

    def embedding_norm(self, x):
        if self.global_l2_norm:
            return x / torch.sqrt(torch.clamp((x * x).sum(), min=1e-12))
        return torch.nn.functional.normalize(x, dim=1, eps=1e-12)

    def towers(self, batch, behaviour=None):
        plan = self._tower_plan()
        names = plan["names"]
        if (self.seq_encoder is None or behaviour is None or torch.is_grad_enabled() and behaviour[0].requires_grad
                or not plan["all_fused"]):
            embs = self.preprocessor.forward_all(batch, names=names)
            return self.towers_from_embeddings(embs, behaviour)
        # One buffer [user features | encoded behaviour sequence | ad features]: the fused bag launch writes the two
        # outer parts in place, the sequence encoder's output drops into the gap, and each tower reads its part as
        # a strided view -- no concatenation copy of the [B, ~1900] tower inputs.
        x, mask = behaviour
        d_seq = self.seq_encoder.d_model
        layout, gap, col = plan["layout"], plan["gap"], plan["total"]
        big = torch.empty(x.shape[0], col, dtype=torch.float32, device=x.device)
        if self.overlap_encoder and x.is_cuda:
            # the sequence encoder (projection GEMM + SDPA, ~0.19 ms at C3) and the fused bag launch (~0.16 ms, DRAM-latency
            # bound at half the HBM bandwidth) are independent until the towers: they run on two streams (a fork / join that
            # CUDA-graph capture records as two branches)
            cur = torch.cuda.current_stream(x.device)
            side = self.__dict__.get("_side_stream")
            if side is None or side.device != x.device:
                side = self.__dict__["_side_stream"] = torch.cuda.Stream(device=x.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                big[:, gap:gap + d_seq] = self.seq_encoder(x, x, x, mask).mean(dim=1)
            for t in (big, x, mask):
                t.record_stream(side)
            self.preprocessor.forward_all(batch, names=names, out=big, layout=layout, views=False)
            cur.wait_stream(side)
        else:
            self.preprocessor.forward_all(batch, names=names, out=big, layout=layout, views=False)
            big[:, gap:gap + d_seq] = self.seq_encoder(x, x, x, mask).mean(dim=1)
        u, a = big[:, :gap + d_seq], big[:, gap + d_seq:]
        if self.global_l2_norm:
            return self.embedding_norm(self.user_dense(u)), self.embedding_norm(self.ad_dense(a))
        return self.user_dense(u, l2_normalize=True), self.ad_dense(a, l2_normalize=True)

    def _tower_plan(self):
        """Feature order, fused-launch eligibility and the column layout [user features | sequence gap | ad features] of the
        towers' input buffer: derived once (228 features make this ~0.3 ms of Python per call otherwise), rebuilt when the
        preprocessing layers change."""
        key = (len(self.preprocessor), tuple(self.user_cols), tuple(self.ad_cols), None if self.seq_encoder is None else self.seq_encoder.d_model)
        plan = self.__dict__.get("_tower_plan_cache")
        if plan is None or plan["key"] != key:
            names = self.user_cols + self.ad_cols
            fused = set(self.preprocessor.fused_names())
            layout, col = {}, 0
            for n in self.user_cols:
                w = self.preprocessor._width(n)
                layout[n] = (col, w)
                col += w
            gap = col
            col += 0 if self.seq_encoder is None else self.seq_encoder.d_model
            for n in self.ad_cols:
                w = self.preprocessor._width(n)
                layout[n] = (col, w)
                col += w
            plan = {"key": key, "names": names, "all_fused": all(n in fused for n in names), "layout": layout, "gap": gap, "total": col}
            self.__dict__["_tower_plan_cache"] = plan
        return plan

    def towers_from_embeddings(self, embs, behaviour=None):
        """The dense part: {feature name: pooled embedding} (+ behaviour sequence) -> normalised tower outputs."""
        user = [embs[n] for n in self.user_cols]
        if self.seq_encoder is not None and behaviour is not None:
            x, mask = behaviour
            user.append(self.seq_encoder(x, x, x, mask).mean(dim=1))
        u = torch.cat(user, dim=-1) if len(user) > 1 else user[0]
        ad = [embs[n] for n in self.ad_cols]
        a = self._adjacent_view(ad) if len(ad) > 1 else ad[0]
        if self.global_l2_norm:
            return self.embedding_norm(self.user_dense(u)), self.embedding_norm(self.ad_dense(a))
        # per-row l2 normalisation rides in the last Dense's epilogue on the tensor-core path
        return self.user_dense(u, l2_normalize=True), self.ad_dense(a, l2_normalize=True)

    def towers_from_fused(self, user_part, ad_part, behaviour=None):
        """The dense part from the two column windows of the fused bag output (training path: differentiable)."""
        u = user_part
        if self.seq_encoder is not None and behaviour is not None:
            x, mask = behaviour
            u = torch.cat([u, self.seq_encoder(x, x, x, mask).mean(dim=1)], dim=-1)
        if self.global_l2_norm:
            return self.embedding_norm(self.user_dense(u)), self.embedding_norm(self.ad_dense(ad_part))
        return self.user_dense(u, l2_normalize=True), self.ad_dense(ad_part, l2_normalize=True)

    @staticmethod
    def _adjacent_view(parts):
        """Column views that sit side by side in one buffer (the fused bag output) are returned as ONE strided view
        instead of being copied by torch.cat; anything else is concatenated."""
        base = parts[0]
        if any(p.requires_grad for p in parts):       # as_strided would cut the graph to parts[1:]
            return torch.cat(parts, dim=-1)
        if all(p.dim() == 2 and p.stride() == base.stride() and p.shape[0] == base.shape[0] for p in parts):
            off, ok = base.storage_offset(), True
            for p in parts:
                ok = ok and p.untyped_storage().data_ptr() == base.untyped_storage().data_ptr() and p.storage_offset() == off
                off += p.shape[1]
            if ok:
                return base.as_strided((base.shape[0], off - base.storage_offset()), base.stride(), base.storage_offset())
        return torch.cat(parts, dim=-1)

    def build(self, device=None):
        """Create every embedding table now (Keras builds variables lazily at the first call): needed before
        load_state_dict() on a fresh model."""
        for layer in self.preprocessor.values():
            if hasattr(layer, "build"):
                layer.build(device)
            elif hasattr(layer, "embedding"):
                layer.embedding.build(device)
        return self

    def forward(self, batch, y_true=None, behaviour=None, training=False):
        u, a = self.towers(batch, behaviour)
        if training:
            return self.loss_fun(y_true, u, a)
        return {"user": u, "ad": a, "label": y_true}
