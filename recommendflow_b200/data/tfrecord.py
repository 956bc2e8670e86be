"""TFRecord / tf.train.Example codec that emits the kernels' input layout directly.

Host-side "next" row of SURVEY.md §8f: the reference writes its training data with
`utils/make_tfrecord.py` (value encoding :26-41, Example building :87-119, GZIP TFRecordWriter
:139-144) and reads it back with `tf.io.parse_example` through `build_feature_description`
(`backend/core/dataloader.py:23-44`).  TensorFlow is not available here, so the wire formats are
implemented from their public specifications:

  TFRecord framing   u64 length | u32 masked_crc32c(length) | bytes | u32 masked_crc32c(bytes),
                     masked = rotr(crc, 15) + 0xa282ead8, the whole file optionally GZIP-ed
  tf.train.Example   Example{1: Features{1: map<string, Feature>}},
                     Feature{oneof 1: BytesList{1: repeated bytes}, 2: FloatList{1: packed float},
                                   3: Int64List{1: packed varint}}

`parse_example` densifies like tf.io.parse_example does for the reference's feature description:
sequence features are padded to the longest list of the BATCH with "" / 0, and string features come
out as a `StringColumn` (arena + offsets), ready for rf_bag_forward -- no Python string objects
on the way to the GPU.
"""
import gzip
import struct
from collections import namedtuple

import numpy as np
import torch

from ..config_parser.config_proto import TYPE_INT, TYPE_STR, FeatureDeal, float32 as TFFloat, int64 as INT64, string as TFString  # noqa: F401 (re-exported dtype names)
from ..strings import StringColumn

# ---- CRC32C (Castagnoli), table driven --------------------------------------------------------
_CRC_TABLE = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ 0x82F63B78 if _c & 1 else _c >> 1
    _CRC_TABLE.append(_c)


def crc32c(data: bytes) -> int:
    crc = 0xFFFFFFFF
    tab = _CRC_TABLE
    for b in data:
        crc = tab[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def masked_crc32c(data: bytes) -> int:
    crc = crc32c(data)
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ---- protobuf primitives ------------------------------------------------------------------------
def _varint(n: int) -> bytes:
    n &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _len_field(field_no: int, payload: bytes) -> bytes:
    return _varint((field_no << 3) | 2) + _varint(len(payload)) + payload


def _read_varint(buf, pos):
    result, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


# ---- value encoding of utils/make_tfrecord.py:26-41 ---------------------------------------------
def _build_int_feature(data):
    """'1,2,3' -> Int64List (make_tfrecord.py:26)."""
    vals = [int(i) for i in str(data).split(",")]
    return _len_field(3, _len_field(1, b"".join(_varint(v) for v in vals)))


def _build_float_feature(data):
    """'0.5,1' -> FloatList (make_tfrecord.py:31)."""
    vals = [float(i) for i in str(data).split(",")]
    return _len_field(2, _len_field(1, struct.pack(f"<{len(vals)}f", *vals)))


def _build_str_feature(data):
    """'a,b' -> BytesList; the missing-value marker "-1" becomes one empty string (make_tfrecord.py:36-41)."""
    data = "" if data == "-1" else str(data)
    return _len_field(1, b"".join(_len_field(1, i.encode()) for i in data.split(",")))


def get_or_ignore_row_data(row, name, na="-1"):
    return row[name] if name in row else na


def build_tfrecord(row, conf):
    """One row (mapping feature name -> TSV cell) -> serialized tf.train.Example (make_tfrecord.py:87-119).
    Like the reference, EVERY feature of the config is written, working or not."""
    entries = []
    for feature in conf.features.features:
        cell = get_or_ignore_row_data(row, feature.name)
        if feature.is_numeric() or feature.is_discrete():
            value = _build_float_feature(cell)
        elif feature.is_hashing() or feature.is_bert_encode():
            value = _build_str_feature(cell)
        elif feature.is_lookup() and feature.py_type == TYPE_INT:
            value = _build_int_feature(cell)
        elif feature.is_lookup() and feature.py_type == TYPE_STR:
            value = _build_str_feature(cell)
        elif feature.is_token_id():
            value = _build_int_feature(cell)
        else:
            raise Exception(f"Unsupported deal method feature: {feature}")
        entry = _len_field(1, feature.name.encode()) + _len_field(2, value)          # map entry {key, value}
        entries.append(_len_field(1, entry))
    return _len_field(1, b"".join(entries))                                           # Example.features


def dump_tfrecord_data(rows, out_file, conf, compression="GZIP"):
    """Rows -> GZIP TFRecord file (make_tfrecord.py:139-144)."""
    opener = gzip.open if compression == "GZIP" else open
    with opener(out_file, "wb") as fh:
        for row in rows:
            rec = build_tfrecord(row, conf)
            head = struct.pack("<Q", len(rec))
            fh.write(head + struct.pack("<I", masked_crc32c(head)) + rec + struct.pack("<I", masked_crc32c(rec)))


def read_tfrecord(path, compression="GZIP", verify_crc=False):
    """Yield the serialized records of a (GZIP) TFRecord file."""
    opener = gzip.open if compression == "GZIP" else open
    with opener(path, "rb") as fh:
        while True:
            head = fh.read(12)
            if not head:
                return
            if len(head) < 12:
                raise IOError("truncated TFRecord header")
            (length,), (hcrc,) = struct.unpack("<Q", head[:8]), struct.unpack("<I", head[8:])
            body = fh.read(length + 4)
            if len(body) < length + 4:
                raise IOError("truncated TFRecord body")
            rec = body[:length]
            if verify_crc:
                if masked_crc32c(head[:8]) != hcrc or masked_crc32c(rec) != struct.unpack("<I", body[length:])[0]:
                    raise IOError("TFRecord CRC mismatch")
            yield rec


def decode_example(rec: bytes):
    """Serialized Example -> {name: ("bytes", [bytes...]) | ("float", [..]) | ("int64", [..])}."""
    out = {}
    pos, end = 0, len(rec)
    while pos < end:
        tag, pos = _read_varint(rec, pos)
        ln, pos = _read_varint(rec, pos)
        if tag != 0x0A:                       # only Example.features (field 1, length-delimited) exists
            pos += ln
            continue
        fpos, fend = pos, pos + ln
        pos = fend
        while fpos < fend:                    # Features.feature map entries
            tag, fpos = _read_varint(rec, fpos)
            ln, fpos = _read_varint(rec, fpos)
            epos, eend = fpos, fpos + ln
            fpos = eend
            key, kind, values = None, None, []
            while epos < eend:                # entry: 1 = key, 2 = Feature
                tag, epos = _read_varint(rec, epos)
                ln, epos = _read_varint(rec, epos)
                if tag == 0x0A:
                    key = rec[epos:epos + ln].decode()
                elif tag == 0x12:
                    vpos, vend = epos, epos + ln
                    while vpos < vend:        # Feature oneof
                        ftag, vpos = _read_varint(rec, vpos)
                        fl, vpos = _read_varint(rec, vpos)
                        lpos, lend = vpos, vpos + fl
                        vpos = lend
                        if ftag == 0x0A:      # BytesList
                            kind = "bytes"
                            while lpos < lend:
                                _, lpos = _read_varint(rec, lpos)
                                bl, lpos = _read_varint(rec, lpos)
                                values.append(rec[lpos:lpos + bl])
                                lpos += bl
                        elif ftag == 0x12:    # FloatList (packed, or repeated fixed32)
                            kind = "float"
                            while lpos < lend:
                                t, lpos = _read_varint(rec, lpos)
                                if t == 0x0A:
                                    pl, lpos = _read_varint(rec, lpos)
                                    values.extend(struct.unpack(f"<{pl // 4}f", rec[lpos:lpos + pl]))
                                    lpos += pl
                                else:
                                    values.append(struct.unpack("<f", rec[lpos:lpos + 4])[0])
                                    lpos += 4
                        elif ftag == 0x1A:    # Int64List (packed, or repeated varint)
                            kind = "int64"
                            while lpos < lend:
                                t, lpos = _read_varint(rec, lpos)
                                if t == 0x0A:
                                    pl, lpos = _read_varint(rec, lpos)
                                    pend = lpos + pl
                                    while lpos < pend:
                                        v, lpos = _read_varint(rec, lpos)
                                        values.append(v - (1 << 64) if v >= (1 << 63) else v)
                                else:
                                    v, lpos = _read_varint(rec, lpos)
                                    values.append(v - (1 << 64) if v >= (1 << 63) else v)
                epos += ln
            if key is not None:
                out[key] = (kind, values)
    return out


# ---- feature description (backend/core/dataloader.py:23-44) and batch parsing ----------------------
FixedLenFeature = namedtuple("FixedLenFeature", "shape dtype default_value")
FixedLenSequenceFeature = namedtuple("FixedLenSequenceFeature", "shape dtype allow_missing default_value")


def build_feature_description(conf):
    desc = {}
    for f in conf.train_features:
        if f.deal == FeatureDeal.Numeric or f.deal == FeatureDeal.Null:
            desc[f.name] = FixedLenFeature((), f.type, f.default)
        elif f.deal in (FeatureDeal.Discrete, FeatureDeal.Hashing, FeatureDeal.Lookup):
            desc[f.name] = FixedLenSequenceFeature((), f.type, True, f.default)
        elif f.deal == FeatureDeal.TokenId:
            desc[f.name] = FixedLenSequenceFeature((), INT64, True, 0)
        elif f.deal == FeatureDeal.BertEncode:
            desc[f.name] = FixedLenFeature((1,), f.type, "")
        elif f.deal in (FeatureDeal.Image, FeatureDeal.Embedding):
            desc[f.name] = FixedLenFeature((), f.type, f.default)
        else:
            raise Exception(f"Unregister Feature: {f.name}")
    return desc


def parse_example(records, feature_description):
    """tf.io.parse_example for the reference's description: a batch of serialized Examples -> dict of dense
    batch tensors.  Sequence features pad to the batch's longest list with the default ("" / 0 / 0.0);
    string features come out as StringColumn [B, Lmax]."""
    decoded = [decode_example(r) for r in records]
    B = len(decoded)
    out = {}
    for name, spec in feature_description.items():
        is_str = spec.dtype.name == "string"
        lists = []
        for ex in decoded:
            kind, vals = ex.get(name, (None, []))
            lists.append(list(vals))
        if isinstance(spec, FixedLenSequenceFeature):
            L = max((len(v) for v in lists), default=0)
            if is_str:
                out[name] = StringColumn.from_lists([v + [b""] * (L - len(v)) for v in lists] if L else [[] for _ in lists])
            else:
                np_dtype = np.int64 if spec.dtype.name == "int64" else np.float32
                arr = np.full((B, L), spec.default_value, dtype=np_dtype)
                for i, v in enumerate(lists):
                    arr[i, :len(v)] = v
                out[name] = torch.from_numpy(arr)
        else:
            if is_str:
                flat = [(v[0] if v else (spec.default_value.encode() if isinstance(spec.default_value, str) else b"")) for v in lists]
                out[name] = StringColumn.from_lists([[x] for x in flat])
            else:
                np_dtype = np.int64 if spec.dtype.name == "int64" else np.float32
                out[name] = torch.from_numpy(np.array([v[0] if v else spec.default_value for v in lists], dtype=np_dtype))
    return out


# ---- native path: librf_b200.so's rf_tfrecord_index / rf_example_parse_columns (include/rf_tfrecord.h) ----------
class RecordFile(object):
    """The records of one (GZIP) TFRecord file, indexed by the native codec: the decompressed stream stays one
    bytes object; `offsets` / `lengths` locate each serialized Example in it."""

    def __init__(self, path, compression="GZIP", verify_crc=False):
        import ctypes as C

        from .. import _native as nat
        opener = gzip.open if compression == "GZIP" else open
        with opener(path, "rb") as fh:
            self.data = np.frombuffer(fh.read(), dtype=np.uint8)
        lib = nat.lib()
        n = C.c_int64(0)
        try:
            nat.check(lib.rf_tfrecord_index(self.data.ctypes.data, self.data.size, 1 if verify_crc else 0, 0, None, None, C.byref(n)))
            self.offsets = np.empty(n.value, dtype=np.int64)
            self.lengths = np.empty(n.value, dtype=np.int64)
            nat.check(lib.rf_tfrecord_index(self.data.ctypes.data, self.data.size, 0, n.value, self.offsets.ctypes.data,
                                            self.lengths.ctypes.data, C.byref(n)))
        except ValueError as e:
            raise IOError(str(e)) from None

    def __len__(self):
        return int(self.offsets.size)


def _pad_rows(counts, L):
    """index [B, L] into the jagged value list: row r's l-th value, clamped to one past its last (a pad)."""
    start = np.zeros(counts.size, dtype=np.int64)
    np.cumsum(counts[:-1], out=start[1:])
    return start[:, None] + np.minimum(np.arange(L, dtype=np.int64)[None, :], counts[:, None].astype(np.int64))


def parse_example_native(rf, first, count, feature_description):
    """tf.io.parse_example over records [first, first + count) of a RecordFile, decoded by the native codec straight
    into arenas / flat arrays (no Python object per value), then densified exactly like `parse_example`."""
    import ctypes as C

    from .. import _native as nat
    names = list(feature_description)
    cols = (nat.ExampleColumn * len(names))()
    keep = []
    for c, name in zip(cols, names):
        spec = feature_description[name]
        raw = name.encode()
        keep.append(raw)
        c.name, c.name_len = raw, len(raw)
        c.kind = {"string": nat.TFR_BYTES, "int64": nat.TFR_INT64}.get(spec.dtype.name, nat.TFR_FLOAT)
        counts = np.empty(count, dtype=np.int32)
        keep.append(counts)
        c.row_counts = counts.ctypes.data
    offs, lens = rf.offsets[first:first + count], rf.lengths[first:first + count]
    lib = nat.lib()
    args = (rf.data.ctypes.data, offs.ctypes.data, lens.ctypes.data, count, cols, len(names))
    nat.check(lib.rf_example_parse_columns(*args, 0))
    outs = []
    for c in cols:
        if c.kind == nat.TFR_BYTES:
            arena = np.zeros(c.n_bytes, dtype=np.uint8)
            voffs = np.empty(c.n_values + 1, dtype=np.int32)
            c.bytes_out, c.value_offsets = arena.ctypes.data, voffs.ctypes.data
            outs.append((arena, voffs))
        elif c.kind == nat.TFR_FLOAT:
            vals = np.empty(c.n_values, dtype=np.float32)
            c.floats_out = vals.ctypes.data
            outs.append((vals,))
        else:
            vals = np.empty(c.n_values, dtype=np.int64)
            c.ints_out = vals.ctypes.data
            outs.append((vals,))
    nat.check(lib.rf_example_parse_columns(*args, 1))
    batch = {}
    for c, name, out in zip(cols, names, outs):
        spec = feature_description[name]
        counts = np.ctypeslib.as_array(C.cast(c.row_counts, C.POINTER(C.c_int32)), shape=(count,)).copy()
        seq = isinstance(spec, FixedLenSequenceFeature)
        L = int(counts.max()) if (seq and count) else 1
        if c.kind == nat.TFR_BYTES:
            arena, voffs = out
            if not seq:                                 # FixedLenFeature: first value, or the default when absent
                default = spec.default_value.encode() if isinstance(spec.default_value, str) else b""
                if (counts == 0).any() and default:
                    strs = [bytes(arena[voffs[i]:voffs[i + 1]]) for i in range(c.n_values)]
                    start = np.concatenate([[0], np.cumsum(counts)[:-1]])
                    batch[name] = StringColumn.from_lists([[strs[start[r]] if counts[r] else default] for r in range(count)])
                    continue
                counts = np.minimum(counts, 1) if (counts <= 1).all() else counts
                idx = _pad_rows(counts, 1) if (counts <= 1).all() else None
                if idx is None:                         # longer lists: keep the first value of each row
                    start = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64)
                    first_end = np.where(counts > 0, voffs[np.minimum(start + 1, c.n_values)], voffs[np.minimum(start, c.n_values)])
                    strs = [bytes(arena[voffs[min(start[r], c.n_values)]:first_end[r]]) for r in range(count)]
                    batch[name] = StringColumn.from_lists([[x] for x in strs])
                    continue
            if L == 0:
                batch[name] = StringColumn.from_lists([[] for _ in range(count)])
                continue
            idx = _pad_rows(counts, L)
            padded = np.empty(count * L + 1, dtype=np.int32)
            padded[:-1] = voffs[idx].ravel()
            padded[-1] = c.n_bytes
            batch[name] = StringColumn.from_arena(arena, padded, (count, L))
        else:
            (vals,) = out
            np_dtype = np.int64 if c.kind == nat.TFR_INT64 else np.float32
            if seq:
                arr = np.full((count, L), spec.default_value, dtype=np_dtype)
                mask = np.arange(L)[None, :] < counts[:, None]
                arr[mask] = vals
                batch[name] = torch.from_numpy(arr)
            else:
                start = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64)
                arr = np.full(count, spec.default_value, dtype=np_dtype)
                has = counts > 0
                arr[has] = vals[start[has]]
                batch[name] = torch.from_numpy(arr)
    return batch


def load_tfrecord(paths, conf, batch_size, compression="GZIP", drop_remainder=False, native=True):
    """Minimal `_get_tfrecord_dataset` (dataloader.py:541-578): batches of (features, labels) dicts.

    native=True decodes with the C codec of librf_b200.so (batches do not span files, as with the reference's
    per-file interleave + batch); native=False is the pure-Python decoder the tests check it against."""
    desc = build_feature_description(conf)
    label_names = conf.features.label_names
    if native:
        for path in ([paths] if isinstance(paths, str) else paths):
            rf = RecordFile(path, compression)
            for first in range(0, len(rf), batch_size):
                count = min(batch_size, len(rf) - first)
                if count < batch_size and drop_remainder:
                    break
                ex = parse_example_native(rf, first, count, desc)
                yield ex, {n: ex[n] for n in label_names if n in ex}
        return
    buf = []
    for path in ([paths] if isinstance(paths, str) else paths):
        for rec in read_tfrecord(path, compression):
            buf.append(rec)
            if len(buf) == batch_size:
                ex = parse_example(buf, desc)
                yield ex, {n: ex[n] for n in label_names if n in ex}
                buf = []
    if buf and not drop_remainder:
        ex = parse_example(buf, desc)
        yield ex, {n: ex[n] for n in label_names if n in ex}


_LIVE_DATASETS = []     # (stop event, worker threads) of iterators that have not been closed yet


def _stop_live_datasets():
    """atexit: worker threads must not run into interpreter shutdown (a daemon thread that is inside torch / numpy
    when the interpreter finalises takes the whole process down with std::terminate)."""
    for stop, threads in list(_LIVE_DATASETS):
        stop.set()
    for stop, threads in list(_LIVE_DATASETS):
        for t in threads:
            t.join(timeout=10)


import atexit  # noqa: E402

atexit.register(_stop_live_datasets)


def get_tfrecord_dataset(paths, feature_description, label_names, batch_size, thread_num=4, compression_type="GZIP",
                         prefetch_buffer_size=8, drop_remainder=False, pin_memory=False):
    """`_get_tfrecord_dataset` (dataloader.py:541-578) on the native codec: `thread_num` host threads read, inflate and
    decode files concurrently (zlib and the C decoder both run without the GIL), at most `prefetch_buffer_size` decoded
    batches wait in the queue, and batches come out in file order (batches do not span files).  Yields
    (features, labels) like the reference's parse_example (:77-89).  pin_memory=True page-locks every batch so that
    the `.to(device, non_blocking=True)` of the layers overlaps with compute."""
    import queue
    import threading
    paths = [paths] if isinstance(paths, str) else list(paths)
    assert len(paths) > 0, "Paths must not be empty"
    thread_num = max(1, min(int(thread_num), len(paths)))
    slots = {i: queue.Queue(maxsize=max(1, prefetch_buffer_size // thread_num + 1)) for i in range(len(paths))}
    stop = threading.Event()
    END = object()

    def put(q, item):
        while not stop.is_set():
            try:
                q.put(item, timeout=0.1)
                return True
            except queue.Full:
                continue
        return False

    def work(worker):
        for fi in range(worker, len(paths), thread_num):
            try:
                rf = RecordFile(paths[fi], compression_type)
                for first in range(0, len(rf), batch_size):
                    count = min(batch_size, len(rf) - first)
                    if count < batch_size and drop_remainder:
                        break
                    ex = parse_example_native(rf, first, count, feature_description)
                    if pin_memory:
                        ex = {k: v.pin_memory() for k, v in ex.items()}
                    if not put(slots[fi], ex):
                        return
                put(slots[fi], END)
            except BaseException as e:                     # surfaces in the consumer
                put(slots[fi], e)
                return

    threads = [threading.Thread(target=work, args=(w,), daemon=True) for w in range(thread_num)]
    live, registry = (stop, threads), _LIVE_DATASETS       # local name: module globals are gone during shutdown
    registry.append(live)
    for t in threads:
        t.start()
    try:
        for fi in range(len(paths)):
            while True:
                item = slots[fi].get()
                if item is END:
                    break
                if isinstance(item, BaseException):
                    raise item
                yield item, {n: item[n] for n in label_names if n in item}
    finally:
        stop.set()
        for t in threads:                                  # never leave a worker running into interpreter shutdown
            t.join(timeout=10)
        if live in registry:
            registry.remove(live)
