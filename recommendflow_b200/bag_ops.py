"""Host-side launcher of the fused hash + gather + pool kernel (rf_bag_forward).

PyTorch is plumbing here: it owns device memory and the current CUDA stream; every computation
happens in librf_b200.so.  A `FieldCall` is one feature field of one batch; `bag_forward`
sends any number of them through ONE kernel launch.
"""
import ctypes as C

import torch

from . import _native as nat
from .strings import StringColumn


class FieldCall(object):
    """Inputs, tables and output slot of one field for one launch.

    keys      StringColumn | int64 tensor [B, L] (hashed as decimal strings) | None
    ids       int64 tensor [T, n_items] of pre-hashed row ids (instead of keys)
    tables    list of (weights[N, D] fp32 cuda tensor or None, num_bins, salt)
    out       fp32 cuda tensor view [B, T*D] (row stride free, columns contiguous) or None
    ids_out   int64 tensor [T, n_items] or None
    """

    def __init__(self, tables, dim, combiner="sum", keys=None, ids=None, mask_mode=nat.MASK_NONE,
                 int_mask_value=0, out=None, ids_out=None, bag_len=None, bag_offsets=None, n_items=None, flags=0,
                 bag_ends=None, mask_bytes=b""):
        self.tables, self.dim, self.combiner = tables, dim, combiner
        self.keys, self.ids = keys, ids
        self.mask_mode, self.int_mask_value = mask_mode, int_mask_value
        self.out, self.ids_out = out, ids_out
        self.bag_len, self.bag_offsets, self.n_items = bag_len, bag_offsets, n_items
        self.flags = flags
        self.bag_ends = bag_ends
        self.mask_bytes = mask_bytes      # MASK_STRING_VALUE: the mask string (Keras Hashing(mask_value="..."))


def _require_cuda(t, what):
    if not t.is_cuda:
        raise nat.NativeError(f"{what} must live on a CUDA device (got {t.device}); there is no CPU fallback")
    return t


def _fill(desc, call, batch):
    keep = []
    if call.combiner not in nat.COMBINER:
        raise ValueError(f"Do not support combiner = '{call.combiner}', supported: [null, sum, min, max, avg, first, last]")
    if len(call.tables) < 1 or len(call.tables) > nat.MAX_TABLES:
        raise ValueError(f"a field takes 1..{nat.MAX_TABLES} tables")
    bag_offsets = call.bag_offsets
    if isinstance(call.keys, StringColumn):
        col = call.keys
        _require_cuda(col.data, "string arena")
        desc.bytes = col.data.data_ptr()
        desc.str_offsets = col.offsets.data_ptr()
        keep += [col.data, col.offsets]
        n_items = col.n_items
        if bag_offsets is None:
            bag_offsets = col.bag_offsets
        bag_len = call.bag_len if call.bag_len is not None else col.shape[1]
    elif call.keys is not None:
        vals = _require_cuda(call.keys, "integer keys")
        if vals.dtype != torch.int64:
            raise ValueError(f"integer keys must be int64, got {vals.dtype}")
        vals = vals.contiguous()
        desc.int_values = vals.data_ptr()
        keep.append(vals)
        n_items = vals.numel()
        bag_len = call.bag_len if call.bag_len is not None else (vals.shape[1] if vals.dim() == 2 else 1)
    elif call.ids is not None:
        ids = _require_cuda(call.ids, "ids")
        if ids.dtype != torch.int64 or not ids.is_contiguous():
            raise ValueError("ids must be a contiguous int64 tensor [n_tables, n_items]")
        desc.ids = ids.data_ptr()
        keep.append(ids)
        # n_items may be an estimate when the true count only exists on the device (sharded path):
        # with one table it only steers tile sizing
        n_items = call.n_items if call.n_items is not None else ids.numel() // len(call.tables)
        bag_len = call.bag_len
    else:
        raise ValueError("a field needs keys or ids")
    if bag_offsets is not None:
        _require_cuda(bag_offsets, "bag_offsets")
        ends = call.bag_ends
        if bag_offsets.dtype != torch.int32 or bag_offsets.numel() < batch + (0 if ends is not None else 1):
            raise ValueError("bag_offsets must be int32 [batch + 1] (or [batch] together with bag_ends)")
        desc.bag_offsets = bag_offsets.data_ptr()
        if ends is not None:
            _require_cuda(ends, "bag_ends")
            if ends.dtype != torch.int32 or ends.numel() < batch:
                raise ValueError("bag_ends must be int32 [batch]")
            desc.bag_ends = ends.data_ptr()
            keep.append(ends)
        desc.n_items = n_items
        keep.append(bag_offsets)
        desc.bag_len = 0
    else:
        if bag_len is None or batch * bag_len != n_items:
            raise ValueError(f"dense field: batch ({batch}) x bag_len ({bag_len}) != n_items ({n_items})")
        desc.bag_len = bag_len
        desc.n_items = n_items
    desc.n_tables = len(call.tables)
    for t, (w, num_bins, salt) in enumerate(call.tables):
        if num_bins is None or num_bins <= 0:
            raise ValueError("`num_bins` cannot be `None` or non-positive values.")
        strong, k0, k1 = nat.salt_to_key(salt)
        td = desc.tables[t]
        td.num_bins, td.use_strong, td.key0, td.key1 = int(num_bins), strong, k0, k1
        if w is not None:
            _require_cuda(w, "embedding table")
            if w.dtype != torch.float32 or not w.is_contiguous() or w.shape != (num_bins, call.dim):
                raise ValueError(f"table {t} must be contiguous fp32 [{num_bins}, {call.dim}], got {tuple(w.shape)} {w.dtype}")
            td.weights = w.data_ptr()
            keep.append(w)
    desc.dim = call.dim
    desc.combiner = nat.COMBINER[call.combiner]
    desc.mask_mode = call.mask_mode
    if call.mask_mode == nat.MASK_STRING_VALUE:
        raw = bytes(call.mask_bytes)
        if len(raw) > nat.MAX_MASK_BYTES:
            raise NotImplementedError(f"a string mask_value takes at most {nat.MAX_MASK_BYTES} bytes")
        C.memmove(desc.mask_bytes, raw, len(raw))
        desc.mask_len = len(raw)
    desc.flags = call.flags
    desc.int_mask_value = int(call.int_mask_value)
    if call.dim > 0:
        out = _require_cuda(call.out, "output")
        if out.dtype != torch.float32 or out.dim() != 2 or out.shape[0] != batch or \
                out.shape[1] != call.dim * len(call.tables) or (out.shape[1] > 1 and out.stride(1) != 1):
            raise ValueError(f"output must be fp32 [batch, n_tables*dim] with contiguous columns, got {tuple(out.shape)}")
        desc.out = out.data_ptr()
        desc.out_stride = out.stride(0) if batch > 1 else out.shape[1]
        keep.append(out)
    if call.ids_out is not None:
        io = _require_cuda(call.ids_out, "ids_out")
        if io.dtype != torch.int64 or not io.is_contiguous() or io.numel() != len(call.tables) * n_items:
            raise ValueError("ids_out must be contiguous int64 [n_tables, n_items]")
        desc.ids_out = io.data_ptr()
        keep.append(io)
    return keep


class BagPlan(object):
    """The C descriptors of one fused launch, built once.  Re-launching a plan costs one C call: use it
    when the same buffers are refilled every step (static key / output buffers), e.g. with hundreds of
    fields where building the descriptors in Python would dominate the step."""

    def __init__(self, calls, batch):
        self.batch = batch
        self.n = len(calls)
        self.descs = (nat.FieldDesc * max(self.n, 1))()
        self.keep = []
        for i, call in enumerate(calls):
            self.keep += _fill(self.descs[i], call, batch)
        self.device = self.keep[0].device if self.keep else None
        if any(t.device != self.device for t in self.keep):
            raise ValueError("all tensors of one launch must be on the same device")

    def launch(self, stream=None, max_ctas_per_sm=0):
        """max_ctas_per_sm > 0 caps this launch's grid (the kernel walks its tiles grid-stride), leaving room on
        every SM for kernels of other streams."""
        if self.n == 0 or self.batch == 0:
            return
        if stream is None:
            stream = torch.cuda.current_stream(self.device)
        with nat.on_device(self.device):
            nat.check(nat.lib().rf_bag_forward_ex(self.descs, self.n, self.batch, int(max_ctas_per_sm),
                                                  C.c_void_p(stream.cuda_stream)))


def bag_forward(calls, batch, stream=None, max_ctas_per_sm=0):
    """Run every FieldCall of one batch through a single rf_bag_forward launch."""
    if not calls or batch == 0:
        return
    plan = BagPlan(calls, batch)
    plan.launch(stream, max_ctas_per_sm)
    return plan.keep


def hash_strings(col, num_bins, mask_value=None, salt=None):
    """Keras `Hashing(num_bins, mask_value, salt)` on a StringColumn -> int64 ids shaped like it."""
    mode, raw = nat.string_mask(mask_value)
    _require_cuda(col.data, "string arena")
    out = torch.empty(col.n_items, dtype=torch.int64, device=col.device)
    strong, k0, k1 = nat.salt_to_key(salt)
    if num_bins is None or num_bins <= 0:
        raise ValueError("`num_bins` cannot be `None` or non-positive values.")
    if col.n_items:
        stream = C.c_void_p(torch.cuda.current_stream(col.device).cuda_stream)
        with nat.on_device(col.device):
            if mode == nat.MASK_STRING_VALUE:     # any other string: the kernel compares the key bytes with it
                nat.check(nat.lib().rf_hash_strings_masked(col.data.data_ptr(), col.offsets.data_ptr(), col.n_items,
                                                           int(num_bins), raw, len(raw), strong, k0, k1, out.data_ptr(), stream))
            else:
                nat.check(nat.lib().rf_hash_strings(col.data.data_ptr(), col.offsets.data_ptr(), col.n_items, int(num_bins),
                                                    mode, strong, k0, k1, out.data_ptr(), stream))
    return out.view(col.shape) if col.shape[1] is not None else out


def hash_ints(values, num_bins, mask_value=None, salt=None):
    """Keras `Hashing` on an int64 tensor (values are hashed as their decimal strings)."""
    _require_cuda(values, "integer keys")
    if values.dtype != torch.int64:
        raise ValueError(f"integer keys must be int64, got {values.dtype}")
    if num_bins is None or num_bins <= 0:
        raise ValueError("`num_bins` cannot be `None` or non-positive values.")
    vals = values.contiguous()
    out = torch.empty_like(vals)
    strong, k0, k1 = nat.salt_to_key(salt)
    mode = nat.MASK_NONE if mask_value is None else nat.MASK_INT_VALUE
    if vals.numel():
        with nat.on_device(vals.device):
            nat.check(nat.lib().rf_hash_int64(vals.data_ptr(), vals.numel(), int(num_bins), mode,
                                              int(mask_value) if mask_value is not None else 0, strong, k0, k1,
                                              out.data_ptr(),
                                              C.c_void_p(torch.cuda.current_stream(vals.device).cuda_stream)))
    return out


def bag_backward(ids, table, grad_out, alpha, combiner="sum", bag_len=None, bag_offsets=None):
    """table[ids[k]] += alpha * grad_out[bag(k)] (x 1/count for "avg"), in place.  `ids`: the int64 bucket ids
    of ONE table as the forward wrote them (ids_out[t]); alpha = -lr fuses the SGD step."""
    ids = _require_cuda(ids, "ids").contiguous().view(-1)
    _require_cuda(table, "table")
    g = _require_cuda(grad_out, "grad_out")
    if g.dtype != torch.float32 or g.dim() != 2 or g.stride(1) != 1 or g.shape[1] != table.shape[1]:
        raise ValueError("grad_out must be fp32 [batch, dim] with contiguous columns")
    if table.dtype != torch.float32 or not table.is_contiguous():
        raise ValueError("table must be contiguous fp32")
    batch = g.shape[0]
    with nat.on_device(table.device):
        nat.check(nat.lib().rf_bag_backward(ids.data_ptr(), ids.numel(), None if bag_offsets is None else bag_offsets.data_ptr(),
                                            bag_len or 0, batch, g.data_ptr(), g.stride(0), table.shape[1],
                                            nat.COMBINER[combiner], float(alpha), table.data_ptr(),
                                            C.c_void_p(torch.cuda.current_stream(table.device).cuda_stream)))
    return table


def bag_minmax_key_grads(ids, table, pooled, grad_out, bag_len=None, bag_offsets=None):
    """Per-key gradient rows [n_keys, dim] of "min" / "max" pooling (TensorFlow's _MinOrMaxGrad: the pooled element's gradient
    goes to the keys whose row element equals it, split equally among ties).  `pooled`: the forward's output for these bags.
    Apply with `bag_backward(ids, table, key_grads, alpha, "sum", bag_len=1)` or `BagAdam.apply(ids, key_grads, "sum", bag_len=1)`."""
    ids = _require_cuda(ids, "ids").contiguous().view(-1)
    _require_cuda(table, "table")
    y, g = _require_cuda(pooled, "pooled"), _require_cuda(grad_out, "grad_out")
    for t, what in ((y, "pooled"), (g, "grad_out")):
        if t.dtype != torch.float32 or t.dim() != 2 or t.stride(1) != 1 or t.shape[1] != table.shape[1]:
            raise ValueError(f"{what} must be fp32 [batch, dim] with contiguous columns")
    if y.shape[0] != g.shape[0]:
        raise ValueError("pooled and grad_out must have one row per bag")
    out = torch.empty(ids.numel(), table.shape[1], dtype=torch.float32, device=table.device)
    with nat.on_device(table.device):
        nat.check(nat.lib().rf_bag_minmax_key_grads(ids.data_ptr(), ids.numel(), None if bag_offsets is None else bag_offsets.data_ptr(),
                                                    bag_len or 0, g.shape[0], table.data_ptr(), table.shape[1], y.data_ptr(), y.stride(0),
                                                    g.data_ptr(), g.stride(0), out.data_ptr(),
                                                    C.c_void_p(torch.cuda.current_stream(table.device).cuda_stream)))
    return out


class BagAdam(object):
    """tf.keras.optimizers.Adam for one embedding table, applied from the pooled bag's gradient
    (rf_bag_backward_adam).  Reference: /root/reference/example/ranking_search/train.py:97-104.

    Keras semantics by default: every row decays its moments and moves each step (`lazy=False`); `lazy=True`
    updates only the gathered rows.  Owns the moment buffers `m`, `v` and the sort workspace."""

    def __init__(self, table, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7, lazy=False):
        _require_cuda(table, "table")
        if table.dtype != torch.float32 or not table.is_contiguous() or table.dim() != 2:
            raise ValueError("table must be a contiguous fp32 [rows, dim] tensor")
        self.table = table
        self.m = torch.zeros_like(table)
        self.v = torch.zeros_like(table)
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon, self.lazy = learning_rate, beta_1, beta_2, epsilon, lazy
        self.iterations = 0
        self._ws = None

    def state_dict(self):
        return {"iterations": self.iterations, "m": self.m.clone(), "v": self.v.clone()}

    def load_state_dict(self, state):
        self.iterations = int(state["iterations"])
        self.m.copy_(state["m"])
        self.v.copy_(state["v"])

    def _workspace(self, n_keys):
        need = int(nat.lib().rf_bag_adam_workspace_bytes(n_keys, self.table.shape[0]))
        if need < 0:
            nat.check(nat.RF_ERR_INVALID)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.table.device)
        return self._ws

    def apply(self, ids, grad_out, combiner="sum", bag_len=None, bag_offsets=None):
        """One optimizer step.  ids: int64 bucket ids of this table as the forward wrote them (ids_out[t]);
        grad_out: fp32 [batch, dim] gradient of the pooled output."""
        ids = _require_cuda(ids, "ids").contiguous().view(-1)
        g = _require_cuda(grad_out, "grad_out")
        dim = self.table.shape[1]
        if g.dtype != torch.float32 or g.dim() != 2 or g.stride(1) != 1 or g.shape[1] != dim:
            raise ValueError("grad_out must be fp32 [batch, dim] with contiguous columns")
        self.iterations += 1
        p = nat.AdamParams(lr=self.learning_rate, beta1=self.beta_1, beta2=self.beta_2, epsilon=self.epsilon,
                           step=self.iterations, lazy=1 if self.lazy else 0)
        with nat.on_device(self.table.device):
            ws = self._workspace(ids.numel())
            nat.check(nat.lib().rf_bag_backward_adam(
                ids.data_ptr(), ids.numel(), None if bag_offsets is None else bag_offsets.data_ptr(), bag_len or 0, g.shape[0],
                g.data_ptr(), g.stride(0), dim, nat.COMBINER[combiner], C.byref(p), self.table.data_ptr(), self.m.data_ptr(),
                self.v.data_ptr(), self.table.shape[0], ws.data_ptr(), ws.numel(),
                C.c_void_p(torch.cuda.current_stream(self.table.device).cuda_stream)))
        return self.table


class BagAdamGroup(object):
    """tf.keras.optimizers.Adam for a set of embedding tables that share `dim`, applied in ONE pass per step
    (rf_bag_backward_adam_multi): one sort / select / update for all tables instead of one per table.

    tables: list of contiguous fp32 [rows_i, dim] CUDA tensors.  Same semantics as `BagAdam` (Keras by default:
    every row of every table decays and moves; `lazy=True`: gathered rows only)."""

    def __init__(self, tables, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7, lazy=False):
        if not tables:
            raise ValueError("BagAdamGroup needs at least one table")
        dim = tables[0].shape[1]
        for t in tables:
            _require_cuda(t, "table")
            if t.dtype != torch.float32 or not t.is_contiguous() or t.dim() != 2 or t.shape[1] != dim:
                raise ValueError("tables must be contiguous fp32 [rows, dim] tensors sharing dim")
        self.tables = list(tables)
        self.m = [torch.zeros_like(t) for t in tables]
        self.v = [torch.zeros_like(t) for t in tables]
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon, self.lazy = learning_rate, beta_1, beta_2, epsilon, lazy
        self.iterations = 0
        self._ws = None

    def state_dict(self):
        return {"iterations": self.iterations, "m": [t.clone() for t in self.m], "v": [t.clone() for t in self.v]}

    def load_state_dict(self, state):
        self.iterations = int(state["iterations"])
        for dst, src in zip(self.m, state["m"]):
            dst.copy_(src)
        for dst, src in zip(self.v, state["v"]):
            dst.copy_(src)
        if getattr(self, "_live", None) is not None:
            self._live.fill_(-1)          # loaded moments may be non-zero anywhere: every row counts as touched (always safe)

    def apply_fused(self, ids, grad, cols, combiners, bag_lens, batch, lr_t=None):
        """The training loop's form of `apply`: table i gathered the keys `ids[i]` (int64 [batch * bag_lens[i]], dense bags)
        and its gradient is the column window grad[:, cols[i] : cols[i] + dim] of ONE [batch, total] gradient tensor.
        The C descriptors are built once and kept: as long as the id buffers stay where they are (the forward's cached
        plan keeps them), a step only re-bases the gradient pointers (one vector addition for all tables) -- the
        round-1 step rebuilt 456 descriptors in Python (~7 ms) every time.
        lr_t: optional 0-dim fp32 CUDA tensor holding lr * sqrt(1 - beta2^t) / (1 - beta1^t); the kernels read it from device
        memory, which is what lets the call be recorded into a CUDA graph (training.GraphedTrainStep)."""
        import numpy as np
        n = len(self.tables)
        if not (len(ids) == len(cols) == len(combiners) == len(bag_lens) == n):
            raise ValueError("one entry per table")
        g = _require_cuda(grad, "grad")
        dim = self.tables[0].shape[1]
        if g.dtype != torch.float32 or g.dim() != 2 or g.stride(1) != 1 or g.shape[0] != batch or g.stride(0) % 4:
            raise ValueError("grad must be fp32 [batch, total] with contiguous columns and a row pitch that is a multiple of 4")
        sig = (tuple(t.data_ptr() for t in ids), tuple(cols), tuple(combiners), tuple(bag_lens), batch, g.stride(0))
        if getattr(self, "_fused_sig", None) != sig:
            arr = (nat.AdamField * n)()
            for i, table in enumerate(self.tables):
                f = arr[i]
                t = _require_cuda(ids[i], "ids")
                if t.dtype != torch.int64 or not t.is_contiguous() or t.numel() != batch * bag_lens[i]:
                    raise ValueError("ids[i] must be contiguous int64 [batch * bag_len]")
                f.table, f.m, f.v = table.data_ptr(), self.m[i].data_ptr(), self.v[i].data_ptr()
                f.table_rows, f.dim = table.shape[0], dim
                f.ids, f.n_keys, f.grad_stride = t.data_ptr(), t.numel(), g.stride(0)
                f.combiner, f.bag_len = nat.COMBINER[combiners[i]], bag_lens[i]
            words = C.sizeof(nat.AdamField) // 8
            self._fused = {"arr": arr, "view": np.frombuffer(arr, dtype=np.uint64).reshape(-1, words),
                           "col": nat.AdamField.grad_out.offset // 8, "cols4": np.asarray(cols, dtype=np.uint64) * np.uint64(4),
                           "ids": list(ids)}
            self._fused_sig = sig
        fz = self._fused
        fz["view"][:, fz["col"]] = np.uint64(g.data_ptr()) + fz["cols4"]
        self.iterations += 1
        dev = self.tables[0].device
        if getattr(self, "_live", None) is None and not self.lazy:
            # rows that ever received a gradient (one bit per row over all tables): the all-rows decay skips the others unread.
            # The moments start at zero; a group restored from a checkpoint marks everything (load_state_dict)
            words = (sum(t.shape[0] for t in self.tables) + 31) // 32 + 1
            fresh = self.iterations == 1 and all(float(m.abs().max()) == 0.0 for m in self.m[:1])
            self._live = torch.zeros(words, dtype=torch.int32, device=dev) if fresh else torch.full((words,), -1, dtype=torch.int32, device=dev)
        p = nat.AdamParams(lr=self.learning_rate, beta1=self.beta_1, beta2=self.beta_2, epsilon=self.epsilon,
                           step=self.iterations, lazy=1 if self.lazy else 0, d_lr_t=None if lr_t is None else lr_t.data_ptr(),
                           d_live_rows=None if self.lazy else self._live.data_ptr())
        with nat.on_device(dev):
            if getattr(self, "_fused_need", None) is None or self._fused_need[0] != sig:
                need = int(nat.lib().rf_bag_adam_multi_workspace_bytes(fz["arr"], n))
                if need < 0:
                    nat.check(nat.RF_ERR_INVALID)
                self._fused_need = (sig, need)
            need = self._fused_need[1]
            if self._ws is None or self._ws.numel() < need:
                self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
            nat.check(nat.lib().rf_bag_backward_adam_multi(fz["arr"], n, batch, C.byref(p), self._ws.data_ptr(), self._ws.numel(),
                                                           C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        return self.tables

    def apply(self, updates, batch):
        """updates: one (ids, grad_out, combiner, bag_len, bag_offsets) per table, in table order; `None` for a
        table that received no gradient this step (it still decays under Keras semantics).  batch: rows of grad_out."""
        if len(updates) != len(self.tables):
            raise ValueError("one update (or None) per table")
        dev = self.tables[0].device
        dim = self.tables[0].shape[1]
        arr = (nat.AdamField * len(self.tables))()
        keep = []
        for i, (table, upd) in enumerate(zip(self.tables, updates)):
            f = arr[i]
            f.table, f.m, f.v = table.data_ptr(), self.m[i].data_ptr(), self.v[i].data_ptr()
            f.table_rows, f.dim = table.shape[0], dim
            f.combiner = nat.COMBINER["sum"]
            if upd is None:
                continue
            ids, g, combiner, bag_len, bag_offsets = upd
            ids = _require_cuda(ids, "ids").contiguous().view(-1)
            g = _require_cuda(g, "grad_out")
            if g.dtype != torch.float32 or g.dim() != 2 or g.stride(1) != 1 or g.shape[1] != dim or g.shape[0] != batch:
                raise ValueError("grad_out must be fp32 [batch, dim] with contiguous columns")
            keep += [ids, g]
            f.ids, f.n_keys, f.grad_out, f.grad_stride = ids.data_ptr(), ids.numel(), g.data_ptr(), g.stride(0)
            f.combiner, f.bag_len = nat.COMBINER[combiner], bag_len or 0
            f.bag_offsets = None if bag_offsets is None else bag_offsets.data_ptr()
        self.iterations += 1
        p = nat.AdamParams(lr=self.learning_rate, beta1=self.beta_1, beta2=self.beta_2, epsilon=self.epsilon,
                           step=self.iterations, lazy=1 if self.lazy else 0)
        with nat.on_device(dev):
            need = int(nat.lib().rf_bag_adam_multi_workspace_bytes(arr, len(self.tables)))
            if need < 0:
                nat.check(nat.RF_ERR_INVALID)
            if self._ws is None or self._ws.numel() < need:
                self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
            nat.check(nat.lib().rf_bag_backward_adam_multi(arr, len(self.tables), batch, C.byref(p), self._ws.data_ptr(),
                                                           self._ws.numel(), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        return self.tables
