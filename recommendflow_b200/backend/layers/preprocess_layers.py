"""Feature-to-embedding layers with the reference's names and constructor surface.

Mirror of /root/reference/backend/layers/preprocess_layers.py -- `EmbeddingBag` (:16-76),
`DoubleHashingEmbedding` (:79-106), `LookupEmbedding` (:135-169), `DiscreteEmbedding`
(:172-200) -- plus `Hashing`, the Keras layer the reference imports (:11) and configures at
:89-90.  Same argument names, same sub-layer names (`<name>_hashing1/2`,
`<name>_embedding_bag1/2`), same `get_config()` keys, same exceptions for bad arguments.
The arithmetic runs in hand-written sm_100a CUDA behind include/rf_b200.h: one fused
hash + gather + pool launch per call instead of the reference's 13 TF ops per feature.

Semantics kept from the reference:
  * inputs are dense [B, L] batches padded with "" / 0; pads are NOT masked out of the pooling:
    a pad hashes to id 0 and row 0 of the table is pooled in; `avg` divides by the padded L;
  * `combiner="null"` returns [B, L, D] and DoubleHashingEmbedding concatenates on axis 1;
  * `first` / `last` index the BATCH axis (`t[0]`, `t[-1]`), as the reference does.
Layers are torch.nn.Modules only so that tables are ordinary parameters (state_dict,
.to(device)); `layer(x)` goes straight to `call`, like the reference's `__call__` override.
"""
import numpy as np
import torch

from ... import _native as nat
from ...bag_ops import FieldCall, bag_forward, hash_ints, hash_strings
from ...config_parser.config_proto import TYPE_INT, TYPE_STR
from ...strings import StringColumn
from ...vocab_ops import DeviceVocabulary, bucketize

SUPPORT_POOLING = ["null", "sum", "min", "max", "avg", "first", "last"]
# bumped whenever any layer's table tensor is (re)created: cached launch plans hold raw table pointers
TABLE_EPOCH = [0]
_POOLED = ("sum", "avg", "min", "max")


def _default_device():
    if not torch.cuda.is_available():
        raise nat.NativeError("recommendflow_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def as_keys(inputs, device=None):
    """Normalise a layer input to what the kernels take: a device StringColumn or an int64 tensor."""
    if isinstance(inputs, StringColumn):
        if inputs.data.is_cuda:
            return inputs
        return inputs.to(device or _default_device(), non_blocking=True)
    if not isinstance(inputs, (torch.Tensor, np.ndarray, list, tuple)) and hasattr(inputs, "__dlpack__"):
        inputs = torch.from_dlpack(inputs)           # zero-copy hand-over from TensorFlow / CuPy / JAX (DLPack)
    if isinstance(inputs, torch.Tensor):
        if inputs.dtype not in (torch.int64, torch.int32):
            raise ValueError(f"tensor inputs must be integer keys, got {inputs.dtype}")
        t = inputs.to(torch.int64)
        if not t.is_cuda:
            t = t.to(device or _default_device(), non_blocking=True)
        return t if t.dim() == 2 else t.reshape(t.shape[0], -1)
    if isinstance(inputs, np.ndarray) and inputs.dtype.kind in "iu":
        return as_keys(torch.from_numpy(inputs.astype(np.int64)), device)
    if isinstance(inputs, np.ndarray):
        return StringColumn.from_numpy(inputs).to(device or _default_device(), non_blocking=True)
    if isinstance(inputs, (list, tuple)):
        flat = inputs[0] if inputs and isinstance(inputs[0], (list, tuple)) else inputs
        if len(flat) and isinstance(flat[0], (int, np.integer)):
            return as_keys(np.asarray(inputs, dtype=np.int64), device)
        return StringColumn.from_lists(inputs).to(device or _default_device(), non_blocking=True)
    raise ValueError(f"unsupported input type {type(inputs).__name__}")


def _batch_and_len(keys):
    if isinstance(keys, StringColumn):
        return keys.shape[0], keys.shape[1]
    return keys.shape[0], keys.shape[1]


class Layer(torch.nn.Module):
    """Minimal stand-in for keras.layers.Layer: a name, get_config(), and call()."""

    def __init__(self, name=None, **kwargs):
        if kwargs:
            raise TypeError(f"unexpected keyword arguments: {sorted(kwargs)}")
        super().__init__()
        self._name = name if name is not None else type(self).__name__.lower()

    @property
    def name(self):
        return self._name

    def get_config(self):
        return {"name": self._name}

    def forward(self, inputs, *args, **kwargs):
        return self.call(inputs, *args, **kwargs)


class Hashing(Layer):
    """Keras `Hashing(num_bins, mask_value=None, salt=None)`: FarmHash64 (no salt) or SipHash-2-4
    (salt) of each value, modulo the bins; with a mask value bin 0 is reserved for it."""

    def __init__(self, num_bins, mask_value=None, salt=None, name=None):
        if num_bins is None or num_bins <= 0:
            raise ValueError(f"The `num_bins` for `Hashing` cannot be `None` or non-positive values. Received: num_bins={num_bins}.")
        super().__init__(name=name)
        nat.salt_to_key(salt)   # validates the salt shape like Keras does
        self.num_bins, self.mask_value, self.salt = num_bins, mask_value, salt

    def call(self, inputs):
        keys = as_keys(inputs)
        if isinstance(keys, StringColumn):
            return hash_strings(keys, self.num_bins, self.mask_value, self.salt)
        return hash_ints(keys, self.num_bins, self.mask_value, self.salt)

    def get_config(self):
        cfg = super().get_config()
        cfg.update({"num_bins": self.num_bins, "salt": self.salt, "mask_value": self.mask_value})
        return cfg


def _new_table(rows, dim, initializer, device):
    w = torch.empty(rows, dim, dtype=torch.float32, device=device)
    if initializer == "uniform":            # Keras 'uniform' == RandomUniform(-0.05, 0.05)
        w.uniform_(-0.05, 0.05)
    elif initializer == "zeros":
        w.zero_()
    elif callable(initializer):
        w.copy_(torch.as_tensor(initializer((rows, dim)), dtype=torch.float32))
    else:
        raise ValueError(f"Unknown initializer: {initializer}")
    return torch.nn.Parameter(w, requires_grad=False)


class EmbeddingBag(Layer):
    """Embedding gather + combiner over axis 1 (pads included)."""

    def __init__(self, input_dim, output_dim, mask_zero=False, combiner="sum", embeddings_initializer="uniform",
                 embeddings_regularizer=None, activity_regularizer=None, embeddings_constraint=None, **kwargs):
        super().__init__(**kwargs)
        if input_dim is None or output_dim is None or input_dim <= 0 or output_dim <= 0:
            raise ValueError(f"Both `input_dim` and `output_dim` should be positive, found input_dim {input_dim} and output_dim {output_dim}")
        self.input_dim, self.output_dim = int(input_dim), int(output_dim)
        self.mask_zero = mask_zero
        self.combiner = combiner
        self.embeddings_initializer = embeddings_initializer
        self.support_pooling = list(SUPPORT_POOLING)
        self.embeddings = None   # created on first use (Keras builds variables lazily too)

    def build(self, device=None):
        if self.embeddings is None:
            self.embeddings = _new_table(self.input_dim, self.output_dim, self.embeddings_initializer,
                                         device or _default_device())
            TABLE_EPOCH[0] += 1
        return self

    def get_weights(self):
        self.build()
        return [self.embeddings.detach().cpu().numpy()]

    def set_weights(self, weights):
        (w,) = weights
        w = torch.as_tensor(np.asarray(w), dtype=torch.float32)
        if tuple(w.shape) != (self.input_dim, self.output_dim):
            raise ValueError(f"Layer {self.name} weight shape {(self.input_dim, self.output_dim)} is not compatible "
                             f"with provided weight shape {tuple(w.shape)}.")
        dev = self.embeddings.device if self.embeddings is not None else _default_device()
        self.embeddings = torch.nn.Parameter(w.to(dev).contiguous(), requires_grad=False)
        TABLE_EPOCH[0] += 1

    def _check_combiner(self):
        if self.combiner not in self.support_pooling:
            raise ValueError(f"Do not support combiner = '{self.combiner}', supported: [{', '.join(self.support_pooling)}]")

    def call(self, inputs, *args, **kwargs):
        self._check_combiner()
        ids = as_keys(inputs)
        if isinstance(ids, StringColumn):
            raise ValueError("EmbeddingBag takes integer ids; hash strings first (DoubleHashingEmbedding)")
        self.build(ids.device)
        B, L = ids.shape
        return _bags_forward([self], self.combiner, B, L, ids=ids.reshape(1, -1).contiguous())

    def get_config(self):
        cfg = super().get_config()
        cfg.update({"combiner": self.combiner})
        return cfg


def _bags_forward(bags, combiner, B, L, keys=None, ids=None, salts=None, mask_mode=nat.MASK_NONE,
                  int_mask_value=0, out=None, mask_bytes=b""):
    """Shared body of EmbeddingBag.call / DoubleHashingEmbedding.call for every combiner."""
    T, D = len(bags), bags[0].output_dim
    dev = bags[0].embeddings.device
    tables = [(b.embeddings.data, b.input_dim, None if salts is None else salts[t]) for t, b in enumerate(bags)]
    if combiner in _POOLED:
        if out is None:
            out = torch.empty(B, T * D, dtype=torch.float32, device=dev)
        if L == 0 or B == 0:
            return out.zero_()
        bag_forward([FieldCall(tables, D, combiner, keys=keys, ids=ids, mask_mode=mask_mode,
                               int_mask_value=int_mask_value, out=out, bag_len=L, mask_bytes=mask_bytes)], B)
        return out
    # null / first / last: a plain gather -- every item is its own bag
    rows = torch.empty(B * L, T * D, dtype=torch.float32, device=dev)
    if B * L:
        bag_forward([FieldCall(tables, D, "sum", keys=keys, ids=ids, mask_mode=mask_mode,
                               int_mask_value=int_mask_value, out=rows, bag_len=1, mask_bytes=mask_bytes)], B * L)
    rows = rows.view(B, L, T, D)
    if combiner == "null":          # concat([E1, E2], axis=1) of two [B, L, D] tensors
        return rows.permute(0, 2, 1, 3).reshape(B, T * L, D)
    picked = rows[0] if combiner == "first" else rows[-1]      # t[0] / t[-1]: the batch axis
    return picked.reshape(L, T * D)                            # concat([L, D], [L, D], axis=1)


def _ids_field_call(bag, ids, out):
    """FieldCall of a single-table bag fed by ready-made ids [B, L] (LookupEmbedding / DiscreteEmbedding)."""
    bag.build(ids.device)
    return FieldCall([(bag.embeddings.data, bag.input_dim, None)], bag.output_dim, bag.combiner,
                     ids=ids.reshape(1, -1).contiguous(), out=out, bag_len=ids.shape[1])


class DoubleHashingEmbedding(Layer):
    """Two salted hashes of the same keys -> two tables -> pooled -> concatenated: [B, 2 * D]."""

    def __init__(self, num_bins, output_dim, seeds, combiner, mask_value=None, mask_zero=False, name=""):
        if num_bins is None or num_bins <= 0:
            raise ValueError("`num_bins` cannot be `None` or non-positive values.")
        super().__init__(name=name)
        self.num_bins = num_bins
        self.output_dim = output_dim
        self.mask_value = mask_value
        self.combiner = combiner
        self.seeds = [seeds, seeds + 7] if isinstance(seeds, int) else seeds
        # The reference indexes the raw `seeds` argument here (preprocess_layers.py:89-90), so an
        # int seed raises TypeError exactly as it does there.
        self.hash1 = Hashing(num_bins, mask_value=mask_value, name=f"{name}_hashing1", salt=seeds[0])
        self.hash2 = Hashing(num_bins, mask_value=mask_value, name=f"{name}_hashing2", salt=seeds[1])
        self.emb1 = EmbeddingBag(num_bins, output_dim, mask_zero, combiner=combiner, name=f"{name}_embedding_bag1")
        self.emb2 = EmbeddingBag(num_bins, output_dim, mask_zero, combiner=combiner, name=f"{name}_embedding_bag2")

    def build(self, device=None):
        self.emb1.build(device)
        self.emb2.build(device)
        return self

    def field_call(self, keys, out):
        """The FieldCall of this layer for a fused multi-field launch (pooled combiners only)."""
        mode, imask, raw = self._mask(keys)
        tables = [(self.emb1.embeddings.data, self.num_bins, self.hash1.salt),
                  (self.emb2.embeddings.data, self.num_bins, self.hash2.salt)]
        return FieldCall(tables, self.output_dim, self.combiner, keys=keys, mask_mode=mode, int_mask_value=imask,
                         out=out, bag_len=_batch_and_len(keys)[1], mask_bytes=raw)

    def _mask(self, keys):
        """(rf_mask_mode, integer mask value, mask string bytes) for these keys."""
        if self.mask_value is None:
            return nat.MASK_NONE, 0, b""
        if isinstance(keys, StringColumn):
            mode, raw = nat.string_mask(self.mask_value)
            return mode, 0, raw
        if isinstance(self.mask_value, str):
            raise ValueError(f"integer keys cannot be compared with the string mask_value {self.mask_value!r}")
        return nat.MASK_INT_VALUE, int(self.mask_value), b""

    def call(self, inputs, *args, **kwargs):
        self.emb1._check_combiner()
        keys = as_keys(inputs)
        self.build(keys.device)
        mode, imask, raw = self._mask(keys)
        B, L = _batch_and_len(keys)
        return _bags_forward([self.emb1, self.emb2], self.combiner, B, L, keys=keys,
                             salts=[self.hash1.salt, self.hash2.salt], mask_mode=mode, int_mask_value=imask, mask_bytes=raw)

    def get_weights(self):
        return self.emb1.get_weights() + self.emb2.get_weights()

    def set_weights(self, weights):
        w1, w2 = weights
        self.emb1.set_weights([w1])
        self.emb2.set_weights([w2])

    def get_config(self):
        cfg = super().get_config()
        cfg.update({"combiner": self.combiner})
        cfg.update({"seeds": self.seeds})
        return cfg


class HashedEmbeddingBag(Layer):
    """Keras `Hashing(num_bins, mask_value, salt)` feeding ONE `EmbeddingBag`: the single-table form of
    DoubleHashingEmbedding (same sub-layer naming, same pad semantics).  The reference only wires the double form
    (preprocess_layers.py:79-106); this is the "hashed sparse field" of BASELINE.json's embedding microbenchmark
    (salt=None: tf.strings.to_hash_bucket_fast = Fingerprint64 mod N) and, with combiner="null", the hashed
    behaviour SEQUENCE whose [B, L, D] embeddings feed the SDPA encoder."""

    def __init__(self, num_bins, output_dim, combiner="sum", salt=None, mask_value=None, mask_zero=False, name=""):
        if num_bins is None or num_bins <= 0:
            raise ValueError("`num_bins` cannot be `None` or non-positive values.")
        super().__init__(name=name)
        self.num_bins, self.output_dim, self.combiner, self.mask_value = num_bins, output_dim, combiner, mask_value
        self.hash = Hashing(num_bins, mask_value=mask_value, name=f"{name}_hashing", salt=salt)
        self.emb = EmbeddingBag(num_bins, output_dim, mask_zero, combiner=combiner, name=f"{name}_embedding_bag")

    def build(self, device=None):
        self.emb.build(device)
        return self

    _mask = DoubleHashingEmbedding._mask

    def field_call(self, keys, out):
        mode, imask, raw = self._mask(keys)
        return FieldCall([(self.emb.embeddings.data, self.num_bins, self.hash.salt)], self.output_dim, self.combiner, keys=keys,
                         mask_mode=mode, int_mask_value=imask, out=out, bag_len=_batch_and_len(keys)[1], mask_bytes=raw)

    def call(self, inputs, *args, out=None, **kwargs):
        self.emb._check_combiner()
        keys = as_keys(inputs)
        self.build(keys.device)
        mode, imask, raw = self._mask(keys)
        B, L = _batch_and_len(keys)
        return _bags_forward([self.emb], self.combiner, B, L, keys=keys, salts=[self.hash.salt], mask_mode=mode,
                             int_mask_value=imask, mask_bytes=raw, out=out)

    def get_weights(self):
        return self.emb.get_weights()

    def set_weights(self, weights):
        self.emb.set_weights(weights)

    def get_config(self):
        cfg = super().get_config()
        cfg.update({"combiner": self.combiner, "num_bins": self.num_bins, "salt": self.hash.salt, "mask_value": self.mask_value})
        return cfg


class LookupEmbedding(Layer):
    """Vocabulary lookup (Keras StringLookup / IntegerLookup: term i -> i + 1, OOV -> 0) + EmbeddingBag
    (/root/reference/backend/layers/preprocess_layers.py:134-168).

    The vocabulary is a device hash table (`vocab_ops.DeviceVocabulary`, rf_vocab_lookup_*); the ids feed
    the fused bag kernel.  The reference's factory passes `vocab_size=len(vocabs)`, one row short of the
    largest index StringLookup can emit; by default the table here has at least
    `len(vocabs) + 1` rows so that the last term never indexes past it.
    reference_rows=True keeps the reference's table shape exactly (`vocab_size` rows, i.e. `len(vocabs)` from its factory) so
    that checkpoints are interchangeable.  The last vocabulary term then has no row; a batch that contains it raises, as the
    reference's CPU gather does ("indices[...] is not in [0, N)") -- checked per call (one device-to-host read of the max id).
    """

    def __init__(self, embedding_dim, dtype, vocabs, vocab_size=None, pooling="sum", name=None, reference_rows=False):
        super().__init__(name=name)
        self.vocabulary = vocabs
        self.pooling = pooling
        self.reference_rows = bool(reference_rows)
        if self.reference_rows:
            vocab_size = int(vocab_size) if vocab_size else len(vocabs)
            if vocab_size < 1:
                raise ValueError("reference_rows needs a non-empty table")
        else:
            vocab_size = max(vocab_size or 0, len(vocabs) + 1)
        if dtype not in (TYPE_STR, TYPE_INT):
            raise ValueError(f"Unsupported type for lookup feature: {dtype}")
        self.key_type = dtype
        self._terms = [int(v) for v in vocabs] if dtype == TYPE_INT else [v if isinstance(v, (str, bytes)) else str(v) for v in vocabs]
        self._vocab = None
        self.embedding = EmbeddingBag(vocab_size, embedding_dim, True, combiner=pooling, name=name + "_embedding")

    def lookup_ids(self, inputs):
        keys = as_keys(inputs)
        if (self.key_type == TYPE_STR) != isinstance(keys, StringColumn):
            raise ValueError(f"lookup feature of type {self.key_type} got {type(keys).__name__} keys")
        device = keys.device
        if self._vocab is None or self._vocab.device != device:
            self._vocab = DeviceVocabulary(self._terms, device)
        ids = self._vocab.lookup(keys)
        if self.reference_rows and len(self._terms) + 1 > self.embedding.input_dim and ids.numel():
            top, rows = int(ids.max()), self.embedding.input_dim
            if top >= rows:
                raise ValueError(f"indices[...] = {top} is not in [0, {rows}): the reference-shaped table "
                                 f"(reference_rows=True) has no row for the last vocabulary term")
        return ids

    def call(self, inputs, *args, **kwargs):
        return self.embedding(self.lookup_ids(inputs))

    def field_call(self, inputs, out):
        """This feature as one field of a fused launch (ids come from the vocabulary kernel)."""
        return _ids_field_call(self.embedding, self.lookup_ids(inputs), out)

    def get_vocabulary(self):
        oov = "[UNK]" if self.key_type == TYPE_STR else -1
        return [oov] + list(self.vocabulary)

    def get_config(self):
        cfg = super().get_config()
        cfg.update({"vocabulary": self.vocabulary})
        cfg.update({"pooling": self.pooling})
        return cfg


class DiscreteEmbedding(Layer):
    """Keras Discretization(bin_boundaries) (bucket = #boundaries <= x) + EmbeddingBag."""

    def __init__(self, embedding_dim, vocabs, vocab_size=None, pooling="sum", name=None):
        super().__init__(name=name)
        self.vocabulary = vocabs
        vocab_size = max(vocab_size or 0, len(vocabs) + 1)
        self.pooling = pooling
        self.bin_boundaries = [float(v) for v in vocabs]
        self._edges = None
        self.embedding = EmbeddingBag(vocab_size, embedding_dim, True, combiner=pooling,
                                      name=name + "_disc_lookup_embedding")

    def bucket_ids(self, inputs):
        x = inputs if isinstance(inputs, torch.Tensor) else torch.as_tensor(np.asarray(inputs, dtype=np.float32))
        if not x.is_cuda:
            x = x.to(_default_device(), non_blocking=True)
        if x.dim() == 1:
            x = x[:, None]
        if self._edges is None or self._edges.device != x.device:
            self._edges = torch.tensor(self.bin_boundaries, dtype=torch.float32).to(x.device)
        return bucketize(x, self._edges)

    def call(self, inputs, *args, **kwargs):
        return self.embedding(self.bucket_ids(inputs))

    def field_call(self, inputs, out):
        return _ids_field_call(self.embedding, self.bucket_ids(inputs), out)

    def get_vocabulary(self):
        return self.bin_boundaries

    def get_config(self):
        cfg = super().get_config()
        cfg.update({"vocabulary": self.vocabulary})
        cfg.update({"pooling": self.pooling})
        return cfg
