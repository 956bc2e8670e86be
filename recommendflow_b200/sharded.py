"""Row-sharded embedding bag: one hashed feature whose table is split across the GPUs of a box.

New design asked for by north_star (the reference only replicates tables under
tf.distribute.MirroredStrategy, /root/reference/backend/utils/gpu_utils.py:13-14).  Row `id`
of the [num_bins, D] table lives on rank `id % world` as local row `id // world`; the batch stays
data-parallel (every rank brings its own bags).  One forward step per rank:

    hash local keys -> route ids to their owners (count / scan / scatter, bag order kept)
      -> owners pool the rows they hold per (source, bag)  -> partial vectors go back
      -> source combines the `world` partials in rank order (avg divides by the bag's key count)

Two transports:
  * "p2p"  (product path): receive buffers live in symmetric memory; the routing kernel writes
    ids/offsets straight into the owner's buffers and the owner's fused gather+pool kernel writes
    each pooled vector straight into the SOURCE rank's buffer through NVLink peer pointers, so the
    "all-to-all" of partials happens inside the compute kernel, tile by tile; only two stream-
    ordered barriers remain.  No host synchronisation anywhere in the step.
  * "nccl" (baseline): the same kernels on local buffers + torch.distributed all_to_all.
Numerics: partial pools are accumulated in key order per owner and combined in rank order, so the
result differs from the single-GPU sequential sum by fp32 rounding only (tests state the bound).
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _native as nat
from .bag_ops import FieldCall, bag_forward, hash_ints, hash_strings
from .strings import StringColumn


class BucketIds(object):
    """Pre-hashed keys: int64 bucket ids in [0, num_bins) -- dense [B, L] or jagged (flat ids + int32 bag_offsets[B + 1]).
    The "pre-hashed ids" path of SURVEY.md §8(d) C4: routing skips the hash and reads 8 bytes per key."""

    def __init__(self, ids, bag_offsets=None):
        if ids.dtype != torch.int64:
            raise ValueError("bucket ids must be int64")
        if bag_offsets is None and ids.dim() != 2:
            raise ValueError("dense bucket ids must be [B, L]; pass bag_offsets for jagged bags")
        self.ids = ids.contiguous()
        self.bag_offsets = bag_offsets
        B = ids.shape[0] if bag_offsets is None else bag_offsets.numel() - 1
        self.shape = (B, ids.shape[1] if bag_offsets is None else None)

    @property
    def n_items(self):
        return self.ids.numel()

    @property
    def device(self):
        return self.ids.device


class CudaShardOps(object):
    """The compute steps of the sharded forward, on the CUDA kernels of librf_b200.so."""

    def hash(self, keys, num_bins, mask_value, salt):
        if isinstance(keys, StringColumn):
            return hash_strings(keys, num_bins, mask_value, salt).reshape(-1)
        return hash_ints(keys, num_bins, mask_value, salt).reshape(-1)

    def route_keys(self, keys, num_bins, mask_value, salt, ids_ws, bag_offsets, bag_len, batch, world, counts_ws,
                   offs_local, offs_dst_ptrs, rows_dst_ptrs):
        """Hash + route in one pass (string keys); ids_ws receives the bucket ids."""
        if mask_value not in (None, ""):
            raise NotImplementedError("only mask_value in (None, '') is supported for string keys")
        strong, k0, k1 = nat.salt_to_key(salt)
        arr_o = (C.c_void_p * world)(*offs_dst_ptrs) if offs_dst_ptrs is not None else None
        arr_r = (C.c_void_p * world)(*rows_dst_ptrs)
        with torch.cuda.device(keys.device):
            nat.check(nat.lib().rf_shard_route_keys(
                keys.data.data_ptr(), keys.offsets.data_ptr(), int(num_bins),
                nat.MASK_NONE if mask_value is None else nat.MASK_EMPTY_STRING, strong, k0, k1, ids_ws.data_ptr(),
                None if bag_offsets is None else bag_offsets.data_ptr(), bag_len or 0, batch, world, counts_ws.data_ptr(),
                offs_local.data_ptr(), arr_o, arr_r, C.c_void_p(torch.cuda.current_stream(keys.device).cuda_stream)))

    def route_tiles(self, keys, num_bins, mask_value, salt, ids_ws, bag_offsets, bag_len, batch, world,
                    rows_dst_ptrs, begin_dst_ptrs, end_dst_ptrs, max_ctas_per_sm=0):
        """Single-pass routing into the gapped "tile" layout (rf_shard_route_tiles)."""
        arrs = [(C.c_void_p * world)(*p) for p in (rows_dst_ptrs, begin_dst_ptrs, end_dst_ptrs)]
        if isinstance(keys, StringColumn):
            if mask_value not in (None, ""):
                raise NotImplementedError("only mask_value in (None, '') is supported for string keys")
            strong, k0, k1 = nat.salt_to_key(salt)
            args = (keys.data.data_ptr(), keys.offsets.data_ptr(), None, int(num_bins),
                    nat.MASK_NONE if mask_value is None else nat.MASK_EMPTY_STRING, strong, k0, k1,
                    None if ids_ws is None else ids_ws.data_ptr())
            dev = keys.device
        else:                                   # pre-hashed int64 ids
            args = (None, None, keys.data_ptr(), 0, 0, 0, 0, 0, None)
            dev = keys.device
        with torch.cuda.device(dev):
            nat.check(nat.lib().rf_shard_route_tiles_ex(*args, None if bag_offsets is None else bag_offsets.data_ptr(),
                                                        bag_len or 0, batch, world, *arrs, int(max_ctas_per_sm),
                                                        C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))

    def route(self, ids, bag_offsets, bag_len, batch, world, counts_ws, offs_local, offs_dst_ptrs, rows_dst_ptrs):
        arr_o = (C.c_void_p * world)(*offs_dst_ptrs) if offs_dst_ptrs is not None else None
        arr_r = (C.c_void_p * world)(*rows_dst_ptrs)
        with torch.cuda.device(ids.device):
            nat.check(nat.lib().rf_shard_route(ids.data_ptr(), None if bag_offsets is None else bag_offsets.data_ptr(),
                                               bag_len or 0, batch, world, counts_ws.data_ptr(), offs_local.data_ptr(),
                                               arr_o, arr_r, C.c_void_p(torch.cuda.current_stream(ids.device).cuda_stream)))

    def pool(self, shard, rows_per_src, offs_per_src, outs_per_src, batch, combiner, est_items, ends_per_src=None,
             accumulate=False, max_ctas_per_sm=0):
        """Owner side: pool the rows every source asked for.  accumulate=True adds each pooled vector into the
        source's (zeroed) buffer with red.global.add instead of storing a per-owner partial."""
        ends_per_src = ends_per_src or [None] * len(rows_per_src)
        flags = nat.FIELD_PARTIAL | (nat.FIELD_ACCUMULATE if accumulate else 0)
        calls = [FieldCall([(shard, shard.shape[0], None)], shard.shape[1], combiner, ids=rows.view(1, -1),
                           bag_offsets=offs, bag_ends=ends, out=out, flags=flags, n_items=est_items)
                 for rows, offs, ends, out in zip(rows_per_src, offs_per_src, ends_per_src, outs_per_src)]
        bag_forward(calls, batch, max_ctas_per_sm=max_ctas_per_sm)

    def combine(self, partials, world, batch, dim, combiner, bag_len, bag_offsets, out):
        with torch.cuda.device(out.device):
            nat.check(nat.lib().rf_combine_partials(partials.data_ptr(), world, batch, dim, nat.COMBINER[combiner], bag_len or 0,
                                                    None if bag_offsets is None else bag_offsets.data_ptr(), out.data_ptr(),
                                                    out.stride(0), C.c_void_p(torch.cuda.current_stream(out.device).cuda_stream)))


    def adam(self, shard, m, v, rows, offs_all, grads, state, params):
        """Keras Adam on this rank's shard from the routed rows of every source (rf_bag_backward_adam):
        rows [n] int64 local rows, offs_all [n_bags + 1] CSR over all (source, bag) pairs, grads [n_bags, D]."""
        n_bags = grads.shape[0]
        live = None
        if not params["lazy"]:
            # rows of this shard that ever received a gradient: the all-rows decay of Keras' Adam skips the others unread
            # (rf_adam_params.d_live_rows; bit-identical).  A state that does not start at step 1 (restored moments) marks all rows.
            if state.get("live") is None:
                words = (shard.shape[0] + 31) // 32 + 1
                state["live"] = (torch.zeros(words, dtype=torch.int32, device=shard.device) if params["step"] == 1 else
                                 torch.full((words,), -1, dtype=torch.int32, device=shard.device))
            live = state["live"].data_ptr()
        p = nat.AdamParams(lr=params["learning_rate"], beta1=params["beta_1"], beta2=params["beta_2"], epsilon=params["epsilon"],
                           step=params["step"], lazy=1 if params["lazy"] else 0, d_live_rows=live)
        with torch.cuda.device(shard.device):
            need = int(nat.lib().rf_bag_adam_workspace_bytes(rows.numel(), shard.shape[0]))
            if need < 0:
                nat.check(nat.RF_ERR_INVALID)
            if state.get("ws") is None or state["ws"].numel() < need:
                state["ws"] = torch.empty(need, dtype=torch.uint8, device=shard.device)
            nat.check(nat.lib().rf_bag_backward_adam(
                rows.data_ptr(), rows.numel(), offs_all.data_ptr(), 0, n_bags, grads.data_ptr(), grads.stride(0), shard.shape[1],
                nat.COMBINER["sum"], C.byref(p), shard.data_ptr(), m.data_ptr(), v.data_ptr(), shard.shape[0],
                state["ws"].data_ptr(), state["ws"].numel(), C.c_void_p(torch.cuda.current_stream(shard.device).cuda_stream)))


class CAbiShardedStep(object):
    """The whole sharded forward step through ONE C call (rf_sharded_bag_forward): what a binder without torch would
    use.  Here torch's symmetric memory only plays the part of "memory every peer can map"; the barriers are the
    library's own (peer-mapped signal pads), no torch.distributed call happens in a step."""

    def __init__(self, num_bins, dim, combiner, max_batch, max_keys, group=None, salt=None, mask_value=""):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.num_bins, self.dim, self.combiner, self.salt, self.mask_value = int(num_bins), int(dim), combiner, salt, mask_value
        dev = torch.device("cuda", torch.cuda.current_device())
        t = torch.tensor([max_keys], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        self.max_batch, self.max_keys = int(max_batch), int(t.item())
        need = int(nat.lib().rf_shard_exchange_bytes(self.world, self.max_batch, self.max_keys, self.dim))
        self.exchange_bytes = (need + 255) // 256 * 256
        self.raw = symm.empty(self.exchange_bytes + 256, dtype=torch.uint8, device=dev)
        self.hdl = symm.rendezvous(self.raw, self.group)
        self.raw.zero_()                                   # the signal pads start at 0
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        self.ctx = nat.ShardCtx(rank=self.rank, world=self.world, max_batch=self.max_batch, max_keys=self.max_keys, dim=self.dim)
        for g, p in enumerate(self.hdl.buffer_ptrs):
            self.ctx.peer_exchange[g] = int(p)
            self.ctx.peer_signals[g] = int(p) + self.exchange_bytes
        self.step = 0

    def __call__(self, keys, shard, out=None):
        if isinstance(keys, StringColumn):
            mode, _ = nat.string_mask(self.mask_value)
            strong, k0, k1 = nat.salt_to_key(self.salt)
            args = (keys.data.data_ptr(), keys.offsets.data_ptr(), None, self.num_bins, mode, strong, k0, k1)
        else:                                              # BucketIds
            args = (None, None, keys.ids.data_ptr(), 0, 0, 0, 0, 0)
        B, L = keys.shape
        if out is None:
            out = torch.empty(B, self.dim, dtype=torch.float32, device=shard.device)
        self.step += 1
        with torch.cuda.device(shard.device):
            nat.check(nat.lib().rf_sharded_bag_forward(
                C.byref(self.ctx), *args, None if keys.bag_offsets is None else keys.bag_offsets.data_ptr(), L or 0, B,
                shard.data_ptr(), shard.shape[0], nat.COMBINER[self.combiner], self.step, out.data_ptr(), out.stride(0),
                C.c_void_p(torch.cuda.current_stream(shard.device).cuda_stream)))
        return out


def shard_rows(num_bins, rank, world):
    """Number of table rows rank `rank` holds (ids rank, rank + world, ...)."""
    return (num_bins - rank + world - 1) // world if num_bins > rank else 0


class ShardedEmbeddingBag(torch.nn.Module):
    def __init__(self, num_bins, output_dim, combiner="sum", salt=None, mask_value="", group=None, transport="p2p",
                 max_batch=8192, max_keys=None, device=None, name="sharded_bag", ops=None, deterministic=True,
                 keep_ids=True):
        super().__init__()
        if num_bins is None or num_bins <= 0:
            raise ValueError("`num_bins` cannot be `None` or non-positive values.")
        if combiner not in ("sum", "avg", "min", "max"):
            raise ValueError(f"Do not support combiner = '{combiner}', supported: [sum, min, max, avg]")
        if transport not in ("p2p", "nccl"):
            raise ValueError("transport must be 'p2p' or 'nccl'")
        self._name = name
        self.num_bins, self.output_dim, self.combiner = int(num_bins), int(output_dim), combiner
        self.salt, self.mask_value = salt, mask_value
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        if self.world > 16:
            raise NotImplementedError("at most 16 ranks (one NVSwitch box)")
        self.transport = transport
        # deterministic=False (p2p, sum / avg): the owners ADD their partial pools straight into the source rank's
        # zeroed buffer over NVLink (red.global.add) -- no per-owner partial buffers, no combine pass; the price is
        # a summation order across owners that changes from run to run (fp32 re-association, same bound as below)
        self.deterministic = bool(deterministic)
        if not self.deterministic and (transport != "p2p" or combiner not in ("sum", "avg")):
            raise ValueError("deterministic=False needs the p2p transport and a sum / avg combiner")
        self.keep_ids = bool(keep_ids)      # string keys: also leave the bucket ids in a workspace (the backward's input)
        self.max_batch = int(max_batch)
        self.max_keys = int(max_keys if max_keys is not None else max_batch * 200)
        self.ops = ops if ops is not None else CudaShardOps()
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        rows = shard_rows(self.num_bins, self.rank, self.world)
        self.shard = torch.nn.Parameter(torch.empty(max(rows, 1), self.output_dim, dtype=torch.float32,
                                                    device=self.device).uniform_(-0.05, 0.05), requires_grad=False)
        self._bufs = None
        # pipelined mode: resident CTAs per SM the pooling kernel may take (of 4), so that the routing
        # kernels of the next step find room on every SM; 0 = no cap
        import os
        self.pool_ctas_per_sm = int(os.environ.get("RF_SHARD_POOL_CTAS", "0"))
        # ... and the routing kernel's own residency cap while it runs under the previous step's pooling: measured on the
        # one-GPU emulation of the 8-way step (tools/emu_sharded.py, profiles/r2_emu_sharded.json): uncapped routing +
        # uncapped pooling 0.761 ms, routing at 2 CTAs/SM 0.736, at 1 CTA/SM 0.750; capping the POOLING kernel instead
        # (3 of 4 CTAs/SM, round 1's default) costs 0.09 ms because its tiles are then walked statically
        self.route_ctas_per_sm = int(os.environ.get("RF_SHARD_ROUTE_CTAS", "2"))
        # p2p routing layout: "tiles" = single-pass routing into gapped per-bag [begin, end) runs;
        # "csr" = count / scan / scatter into a gap-free CSR (what the nccl transport always uses)
        self.route_layout = os.environ.get("RF_SHARD_ROUTE", "tiles")
        self.profile = None      # set to [] to collect (phase name, cuda event) pairs per forward

    def _tick(self, phase):
        if self.profile is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.profile.append((phase, ev))

    @property
    def name(self):
        return self._name

    # ---- weights ---------------------------------------------------------------------------------
    def set_full_weights(self, full):
        """Load this rank's rows (rank::world) of a full [num_bins, D] table."""
        full = torch.as_tensor(full, dtype=torch.float32)
        if tuple(full.shape) != (self.num_bins, self.output_dim):
            raise ValueError(f"expected a [{self.num_bins}, {self.output_dim}] table, got {tuple(full.shape)}")
        mine = full[self.rank::self.world].contiguous()
        if mine.shape[0] == 0:
            mine = torch.zeros(1, self.output_dim)
        self.shard = torch.nn.Parameter(mine.to(self.device), requires_grad=False)

    # ---- buffers ---------------------------------------------------------------------------------
    N_SETS = 2      # buffer sets: route(i+1), pool(i) and the NVLink drain + combine of (i-1) overlap

    def _alloc(self):
        dev = self.device
        if self.world > 1:       # symmetric buffers must have one size on every rank
            t = torch.tensor([self.max_keys], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            self.max_keys = int(t.item())
        W, B, K, D = self.world, self.max_batch, self.max_keys, self.output_dim
        b = {"counts": torch.empty(W * B, dtype=torch.int32, device=dev),
             "offs_local": torch.empty(W * (B + 1 + (B + 1023) // 1024), dtype=torch.int32, device=dev),
             "ids_ws": None, "step": 0, "rows_free": [None] * self.N_SETS, "combined": [None] * self.N_SETS}
        if self.transport == "p2p":
            import torch.distributed._symmetric_memory as symm
            rows_bytes = W * K * 8
            offs_bytes = (2 * W * (B + 1) * 4 + 15) // 16 * 16      # [2][W][B+1]: bag begins / ends (or one CSR)
            part_bytes = (W if self.deterministic else 1) * B * D * 4
            set_bytes = rows_bytes + offs_bytes + part_bytes
            raw = symm.empty(self.N_SETS * set_bytes, dtype=torch.uint8, device=dev)
            hdl = symm.rendezvous(raw, self.group)
            b.update(raw=raw, hdl=hdl, set_bytes=set_bytes, rows_bytes=rows_bytes)
            b["rows_recv"], b["offs_recv"], b["ends_recv"], b["partials"], b["peer_partials"] = [], [], [], [], []
            for j in range(self.N_SETS):
                o = j * set_bytes
                b["rows_recv"].append(raw[o:o + rows_bytes].view(torch.int64).view(W, K))
                b["offs_recv"].append(raw[o + rows_bytes:o + rows_bytes + W * (B + 1) * 4].view(torch.int32).view(W, B + 1))
                e0 = o + rows_bytes + W * (B + 1) * 4
                b["ends_recv"].append(raw[e0:e0 + W * (B + 1) * 4].view(torch.int32).view(W, B + 1))
                po = o + rows_bytes + offs_bytes
                if self.deterministic:
                    b["partials"].append(raw[po:po + part_bytes].view(torch.float32).view(W, B, D))
                    b["peer_partials"].append([hdl.get_buffer(r, (W, B, D), torch.float32, po // 4) for r in range(W)])
                else:       # one [B, D] accumulator per set; every owner reduces into it
                    b["partials"].append(raw[po:po + part_bytes].view(torch.float32).view(1, B, D))
                    b["peer_partials"].append([hdl.get_buffer(r, (1, B, D), torch.float32, po // 4) for r in range(W)])
            b["peer_ptr"] = [int(p) for p in hdl.buffer_ptrs]
            # routing and combine are short, latency-bound kernels: give their streams priority so their
            # CTAs slot in between the CTAs of the long HBM-bound pooling kernel instead of queueing behind it
            b["sR"] = torch.cuda.Stream(device=dev, priority=-1)
            b["sP"] = torch.cuda.Stream(device=dev, priority=0)
            b["sC"] = torch.cuda.Stream(device=dev, priority=-1)
        else:
            b["rows_send"] = torch.empty(W, K, dtype=torch.int64, device=dev)
            b["offs_send"] = torch.empty(W, B + 1, dtype=torch.int32, device=dev)
            b["offs_recv"] = [torch.empty(W, B + 1, dtype=torch.int32, device=dev)]
            b["part_send"] = torch.empty(W, B, D, dtype=torch.float32, device=dev)
            b["partials"] = [torch.empty(W, B, D, dtype=torch.float32, device=dev)]
        self._bufs = b
        return b

    # ---- forward ---------------------------------------------------------------------------------
    def _describe(self, keys):
        if isinstance(keys, (StringColumn, BucketIds)):
            B, L, bag_offsets, n_keys = keys.shape[0], keys.shape[1], keys.bag_offsets, keys.n_items
        else:
            B, L, bag_offsets, n_keys = keys.shape[0], keys.shape[1], None, keys.numel()
        if B != self.max_batch:
            raise ValueError(f"batch {B} != max_batch {self.max_batch} the exchange buffers were sized for")
        if n_keys > self.max_keys:
            raise ValueError(f"{n_keys} keys exceed max_keys={self.max_keys}")
        return B, L, bag_offsets, n_keys

    def _route(self, keys, B, L, bag_offsets, offs_dst, rows_dst):
        b, W = self._bufs, self.world
        if isinstance(keys, StringColumn) and hasattr(self.ops, "route_keys"):      # hashing fused into routing
            if b["ids_ws"] is None:
                b["ids_ws"] = torch.empty(self.max_keys, dtype=torch.int64, device=self.device)
            self.ops.route_keys(keys, self.num_bins, self.mask_value, self.salt, b["ids_ws"], bag_offsets, L, B, W,
                                b["counts"], b["offs_local"], offs_dst, rows_dst)
        else:
            ids = self.ops.hash(keys, self.num_bins, self.mask_value, self.salt)
            self._tick("hash")
            self.ops.route(ids, bag_offsets, L, B, W, b["counts"], b["offs_local"], offs_dst, rows_dst)

    def prepare(self, keys, overlap=True, capturing=False):
        """p2p transport, stage 1: hash + route this rank's keys into the owners' receive buffers.

        With overlap=True the three stages of a step run on three streams -- route (+ the "routing has landed"
        barrier) | fused gather+pool into peer memory | NVLink drain + combine -- so that in a loop
        `t = prepare(next); finish(prev)` the latency-bound routing of step i+1, the HBM-bound pooling of step i
        and the NVLink-bound drain of step i-1 proceed concurrently (double-buffered exchange sets); the pooling
        stream then runs pooling kernels back to back (round 1 kept the first barrier on it: 13 us per step).
        capturing=True: the call is being recorded into a CUDA graph together with the `finish` of the previous
        step; dependencies on EARLIER steps are then carried by the order of graph launches, not by events (an
        event recorded in another capture cannot be waited on).  Returns a ticket."""
        if self.transport != "p2p":
            raise ValueError("prepare()/finish() pipelining needs the p2p transport")
        b = self._bufs or self._alloc()
        B, L, bag_offsets, n_keys = self._describe(keys)
        W, K, me = self.world, self.max_keys, self.rank
        j = b["step"] % self.N_SETS
        b["step"] += 1
        base = [b["peer_ptr"][g] + j * b["set_bytes"] for g in range(W)]
        rows_dst = [base[g] + me * K * 8 for g in range(W)]
        offs_dst = [base[g] + b["rows_bytes"] + me * (B + 1) * 4 for g in range(W)]
        ends_dst = [base[g] + b["rows_bytes"] + (W + me) * (B + 1) * 4 for g in range(W)]
        tiles = self.route_layout == "tiles" and hasattr(self.ops, "route_tiles")
        cur = torch.cuda.current_stream(self.device)
        stream = b["sR"] if overlap else cur
        routed = None
        if overlap:
            stream.wait_stream(cur)                       # the keys were produced on the caller's stream
        with torch.cuda.stream(stream):
            if overlap and not capturing and b["rows_free"][j] is not None:
                stream.wait_event(b["rows_free"][j])      # every owner is done pooling out of set j
            self._tick("start")
            if isinstance(keys, BucketIds) and not tiles:
                raise NotImplementedError("pre-hashed BucketIds take the tile routing layout")
            if tiles:
                src = keys
                if isinstance(keys, BucketIds):
                    src = keys.ids.view(-1)
                elif not isinstance(keys, StringColumn):
                    src = self.ops.hash(keys, self.num_bins, self.mask_value, self.salt)
                elif b["ids_ws"] is None and self.keep_ids:
                    b["ids_ws"] = torch.empty(self.max_keys, dtype=torch.int64, device=self.device)
                self.ops.route_tiles(src, self.num_bins, self.mask_value, self.salt, b["ids_ws"], bag_offsets, L, B, W,
                                     rows_dst, offs_dst, ends_dst, max_ctas_per_sm=self.route_ctas_per_sm if overlap else 0)
            else:
                self._route(keys, B, L, bag_offsets, offs_dst, rows_dst)
            if not self.deterministic:
                # the accumulator of this set: its last reader (the finishing pass two steps ago) is ordered before
                # this stream by the rows_free / combined events below; owners only add after the next barrier
                if overlap and not capturing and b["combined"][j] is not None:
                    stream.wait_event(b["combined"][j])
                b["partials"][j].zero_()
            self._tick("route")
            if overlap:
                # after this barrier: every source's routing of this step has landed at its owners, and every rank
                # is past the combine that last read partial set j -- issued HERE, behind the routing, so that it
                # completes under the previous step's pooling instead of delaying this step's
                if not capturing and b["combined"][j] is not None:
                    stream.wait_event(b["combined"][j])
                b["hdl"].barrier(channel=0)
                routed = torch.cuda.Event()
                routed.record(stream)
                if capturing:
                    cur.wait_event(routed)                # join the side stream back into the capturing stream
        return {"set": j, "routed": routed, "B": B, "L": L, "bag_offsets": bag_offsets, "n_keys": n_keys,
                "overlap": overlap, "tiles": tiles, "capturing": capturing}

    def finish(self, ticket, out=None):
        """Stages 2 and 3: barrier, fused gather+pool writing each pooled vector into the SOURCE rank's
        buffer over NVLink, barrier (drain), combine.  The caller's stream waits for the result."""
        b, W, D, me = self._bufs, self.world, self.output_dim, self.rank
        B, L, bag_offsets, j, overlap = ticket["B"], ticket["L"], ticket["bag_offsets"], ticket["set"], ticket["overlap"]
        if out is None:
            out = torch.empty(B, D, dtype=torch.float32, device=self.device)
        cur = torch.cuda.current_stream(self.device)
        hdl = b["hdl"]
        sP = b["sP"] if overlap else cur
        sC = b["sC"] if overlap else cur
        partial_op = "sum" if self.combiner == "avg" else self.combiner
        capturing = ticket.get("capturing", False) or torch.cuda.is_current_stream_capturing()
        with torch.cuda.stream(sP):
            if overlap and capturing:
                sP.wait_stream(cur)                       # fork; the routing + barrier of this ticket ran in an earlier launch
            elif overlap:
                sP.wait_event(ticket["routed"])           # routing landed everywhere (the barrier is behind this event)
            else:
                # after this barrier: every source's routing of this step has landed here, and every rank
                # is past the combine that last read partial set j
                hdl.barrier(channel=0)
            self._tick("barrier0")
            # sources in rotated order (me, me+1, ...): at any moment the W owners write their pooled
            # vectors to W different ranks -- a permutation, not an incast on one rank's NVLink port
            order = [(me + k) % W for k in range(W)]
            acc = not self.deterministic
            self.ops.pool(self.shard.data, [b["rows_recv"][j][s] for s in order],
                          [b["offs_recv"][j][s] for s in order],
                          [b["peer_partials"][j][s][0 if acc else me] for s in order],
                          B, partial_op, max(1, ticket["n_keys"] // W),
                          [b["ends_recv"][j][s] for s in order] if ticket["tiles"] else None,
                          accumulate=acc, max_ctas_per_sm=self.pool_ctas_per_sm if overlap else 0)
            self._tick("pool")
            pooled = None
            if overlap:
                pooled = torch.cuda.Event()
                pooled.record(sP)
        with torch.cuda.stream(sC):
            if overlap:
                sC.wait_event(pooled)
                sC.wait_stream(cur)                       # `out` may still be read by the caller's stream
            hdl.barrier(channel=1)                        # every owner's partials have landed (NVLink drained)
            self._tick("barrier1")
            if overlap and not capturing:
                ev = torch.cuda.Event()
                ev.record(sC)
                b["rows_free"][j] = ev
            # deterministic: reduce the W partials in rank order; else the single accumulator only needs the
            # avg division (or a copy) on its way to `out` -- the same kernel with world = 1
            self.ops.combine(b["partials"][j], W if self.deterministic else 1, B, D, self.combiner, L, bag_offsets, out)
            self._tick("combine")
            if overlap:
                done = torch.cuda.Event()
                done.record(sC)
                if not capturing:
                    b["combined"][j] = done
                cur.wait_event(done)
        # what the backward needs: this step's routing stays in exchange set j until the step after next overwrites it
        self._saved = {"p2p_set": j, "tiles": ticket["tiles"], "B": B, "L": L, "bag_offsets": bag_offsets}
        return out

    def forward(self, keys, out=None):
        """keys: StringColumn (dense [B, L] or jagged), int64 [B, L] tensor (hashed as decimal strings, like Keras
        Hashing), or BucketIds (pre-hashed ids), on this rank's device.
        Every rank must call with the same batch size.  Returns [B, D] fp32."""
        if self.transport == "p2p":
            return self.finish(self.prepare(keys, overlap=False), out)
        b = self._bufs or self._alloc()
        B, L, bag_offsets, n_keys = self._describe(keys)
        W, D = self.world, self.output_dim
        if out is None:
            out = torch.empty(B, D, dtype=torch.float32, device=self.device)
        self._tick("start")
        rows_dst = [b["rows_send"][g].data_ptr() for g in range(W)]
        offs_dst = [b["offs_send"][g].data_ptr() for g in range(W)]
        self._route(keys, B, L, bag_offsets, offs_dst, rows_dst)
        offs_recv = b["offs_recv"][0]
        dist.all_to_all_single(offs_recv, b["offs_send"], group=self.group)
        send_tot = b["offs_send"][:, B].tolist()         # host sync: NCCL needs the split sizes
        recv_tot = offs_recv[:, B].tolist()
        send_flat = torch.cat([b["rows_send"][g, :send_tot[g]] for g in range(W)])
        recv_flat = torch.empty(sum(recv_tot), dtype=torch.int64, device=self.device)
        dist.all_to_all_single(recv_flat, send_flat, recv_tot, send_tot, group=self.group)
        rows_in = list(torch.split(recv_flat, recv_tot))
        # what the backward needs: the routed local rows of every source, in (source, bag, key) order
        self._saved = {"rows": recv_flat, "recv_tot": recv_tot, "offs_recv": offs_recv, "B": B, "L": L, "bag_offsets": bag_offsets}
        partial_op = "sum" if self.combiner == "avg" else self.combiner
        self.ops.pool(self.shard.data, rows_in, [offs_recv[s] for s in range(W)],
                      [b["part_send"][s] for s in range(W)], B, partial_op, max(1, n_keys // W))
        dist.all_to_all_single(b["partials"][0], b["part_send"], group=self.group)
        self.ops.combine(b["partials"][0], W, B, D, self.combiner, L, bag_offsets, out)
        self._tick("combine")
        return out

    def optimizer_state_dict(self):
        """Adam moments of this rank's shard + the iteration counter (the shard itself is `self.shard`, a Parameter)."""
        st = getattr(self, "_adam", None)
        if st is None:
            return {"iterations": 0, "m": None, "v": None}
        return {"iterations": st["iterations"], "m": st["m"].clone(), "v": st["v"].clone()}

    def load_optimizer_state_dict(self, state):
        if state["m"] is None:
            self._adam = None
            return
        self._adam = {"m": state["m"].to(self.shard.device).clone(), "v": state["v"].to(self.shard.device).clone(),
                      "iterations": int(state["iterations"]), "ws": None}

    # ---- backward + optimizer (SURVEY.md §8f rank 1, row-sharded) -------------------------------------
    def apply_adam(self, grad_out, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7, lazy=False):
        """Backward of the LAST forward fused with tf.keras.optimizers.Adam on this rank's rows.

        grad_out: [B, D] gradient of this rank's pooled output.  The reverse exchange is one all-gather of
        the (pre-scaled) gradients: every owner already holds, from the forward's routing, the local rows each
        source gathered per bag, so it can scatter source s's bag gradients to its rows directly.  Rows owned
        by this rank that no source touched decay and move like every Keras Adam variable (lazy=False).
        p2p transport: the sources PUSH their gradients into every owner's receive buffer through the NVLink peer mapping
        (two device-side barriers, no NCCL call), and the owner reads the routing of the forward from the exchange set it
        is still sitting in.  sum / avg combiners."""
        if self.combiner not in ("sum", "avg"):
            raise NotImplementedError(f"backward is implemented for sum / avg pooling, not {self.combiner}")
        sv = getattr(self, "_saved", None)
        if sv is None:
            raise RuntimeError("apply_adam needs a preceding forward")
        B, W, D = sv["B"], self.world, self.output_dim
        g = grad_out.to(torch.float32)
        if tuple(g.shape) != (B, D):
            raise ValueError(f"grad_out must be [{B}, {D}]")
        if self.combiner == "avg":     # the bag's TOTAL key count lives with the source: scale here, owners then sum
            if sv["bag_offsets"] is not None:
                cnt = (sv["bag_offsets"][1:] - sv["bag_offsets"][:-1]).clamp(min=1).to(torch.float32)
                g = g / cnt[:, None]
            else:
                g = g / float(max(sv["L"], 1))
        g = g.contiguous()
        if self.transport == "p2p":
            gathered, rows, offs_all = self._p2p_reverse_exchange(g, sv)
        else:
            gathered = torch.empty(W, B, D, dtype=torch.float32, device=g.device)
            dist.all_gather(list(gathered.unbind(0)), g, group=self.group)
            # one CSR over all (source, bag) pairs: source s's offsets shifted by the keys of the sources before it
            totals = [0]
            for n in sv["recv_tot"]:
                totals.append(totals[-1] + int(n))
            base = torch.tensor(totals, dtype=torch.int64, device=g.device)
            offs_all = (sv["offs_recv"][:, :B].to(torch.int64) + base[:W, None]).reshape(-1)
            offs_all = torch.cat([offs_all, base[W:]]).to(torch.int32).contiguous()
            rows = sv["rows"]
        st = getattr(self, "_adam", None)
        if st is None:
            st = self._adam = {"m": torch.zeros_like(self.shard.data), "v": torch.zeros_like(self.shard.data), "iterations": 0,
                               "ws": None}
        st["iterations"] += 1
        params = {"learning_rate": learning_rate, "beta_1": beta_1, "beta_2": beta_2, "epsilon": epsilon,
                  "step": st["iterations"], "lazy": lazy}
        self.ops.adam(self.shard.data, st["m"], st["v"], rows, offs_all, gathered.view(W * B, D), st, params)
        self._saved = None
        return self.shard

    def _p2p_reverse_exchange(self, g, sv):
        """p2p transport: (gradients of every source [W, B, D], this owner's routed rows in (source, bag, key) order, their CSR
        [W * B + 1]).  Gradients travel by peer stores into a symmetric receive buffer; the routed rows are compacted out of the
        forward's exchange set (the tile route leaves gaps between bags: per-tile capacity slots)."""
        import torch.distributed._symmetric_memory as symm
        b, W, D, me = self._bufs, self.world, self.output_dim, self.rank
        B = sv["B"]
        if b.get("grad_recv") is None:
            raw = symm.empty(W * self.max_batch * D, dtype=torch.float32, device=self.device)
            hdl = symm.rendezvous(raw, self.group)
            b["grad_raw"], b["grad_hdl"] = raw, hdl
            b["grad_recv"] = raw.view(W, self.max_batch, D)
            b["grad_peer"] = [hdl.get_buffer(r, (W, self.max_batch, D), torch.float32, 0) for r in range(W)]
        hdl = b["grad_hdl"]
        hdl.barrier(channel=0)                           # every owner is done with the gradients of the previous step
        for k in range(W):                               # rotated: at any moment the W sources store to W different owners
            r = (me + k) % W
            b["grad_peer"][r][me, :B].copy_(g)
        hdl.barrier(channel=1)                           # every source's gradient has landed here
        gathered = b["grad_recv"][:, :B]
        if B != self.max_batch:
            gathered = gathered.contiguous()
        j = sv["p2p_set"]
        begins = b["offs_recv"][j][:, :B].to(torch.int64)                                  # [W, B]
        ends = (b["ends_recv"][j][:, :B] if sv["tiles"] else b["offs_recv"][j][:, 1:B + 1]).to(torch.int64)
        lens = (ends - begins).reshape(-1)
        offs_all = torch.zeros(W * B + 1, dtype=torch.int64, device=self.device)
        torch.cumsum(lens, dim=0, out=offs_all[1:])
        n = int(offs_all[-1].item())                     # host sync: the sort below is sized by it (the nccl path syncs for its split sizes)
        src_base = (torch.arange(W, device=self.device, dtype=torch.int64) * self.max_keys)[:, None]
        first = (begins + src_base).reshape(-1) - offs_all[:-1]                             # where bag (s, b) starts, minus its compact start
        idx = torch.repeat_interleave(first, lens, output_size=n) + torch.arange(n, device=self.device, dtype=torch.int64)
        rows = b["rows_recv"][j].reshape(-1)[idx]
        return gathered, rows, offs_all.to(torch.int32).contiguous()
