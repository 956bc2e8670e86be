"""ctypes binding of librf_b200.so (the C-ABI in include/rf_b200.h).

There is no fallback: if the shared library is missing or a call fails, this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "librf_b200.so")

RF_OK, RF_ERR_INVALID, RF_ERR_CUDA, RF_ERR_UNSUPPORTED = 0, -1, -2, -3
COMBINER = {"sum": 0, "avg": 1, "min": 2, "max": 3}
MASK_NONE, MASK_EMPTY_STRING, MASK_INT_VALUE, MASK_STRING_VALUE = 0, 1, 2, 3
MAX_MASK_BYTES = 32
ACTIVATION = {None: 0, "linear": 0, "relu": 1, "selu": 2, "tanh": 3, "sigmoid": 4, "gelu": 5}
MAX_TABLES = 2
FIELD_PARTIAL = 1
FIELD_ACCUMULATE = 2


class TableDesc(C.Structure):
    _fields_ = [("weights", C.c_void_p), ("num_bins", C.c_int64), ("use_strong", C.c_int32),
                ("reserved", C.c_int32), ("key0", C.c_uint64), ("key1", C.c_uint64)]


class FieldDesc(C.Structure):
    _fields_ = [("bytes", C.c_void_p), ("str_offsets", C.c_void_p), ("int_values", C.c_void_p),
                ("ids", C.c_void_p), ("bag_offsets", C.c_void_p), ("bag_ends", C.c_void_p), ("n_items", C.c_int64),
                ("bag_len", C.c_int32), ("n_tables", C.c_int32), ("tables", TableDesc * MAX_TABLES),
                ("dim", C.c_int32), ("combiner", C.c_int32), ("mask_mode", C.c_int32), ("flags", C.c_int32),
                ("int_mask_value", C.c_int64), ("out", C.c_void_p), ("out_stride", C.c_int64),
                ("ids_out", C.c_void_p), ("mask_bytes", C.c_uint8 * MAX_MASK_BYTES), ("mask_len", C.c_int32),
                ("reserved", C.c_int32)]


class VocabDesc(C.Structure):
    _fields_ = [("term_bytes", C.c_void_p), ("term_offsets", C.c_void_p), ("term_ints", C.c_void_p),
                ("slots", C.c_void_p), ("capacity", C.c_int64), ("n_terms", C.c_int64)]


class ShardCtx(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("max_batch", C.c_int64), ("max_keys", C.c_int64), ("dim", C.c_int32),
                ("reserved", C.c_int32), ("peer_exchange", C.c_void_p * 16), ("peer_signals", C.c_void_p * 16)]


class AdamParams(C.Structure):
    _fields_ = [("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("epsilon", C.c_float),
                ("step", C.c_int64), ("lazy", C.c_int32), ("reserved", C.c_int32), ("d_lr_t", C.c_void_p),
                ("d_live_rows", C.c_void_p)]


class AdamField(C.Structure):
    _fields_ = [("ids", C.c_void_p), ("bag_offsets", C.c_void_p), ("n_keys", C.c_int64), ("bag_len", C.c_int32),
                ("combiner", C.c_int32), ("grad_out", C.c_void_p), ("grad_stride", C.c_int64), ("table", C.c_void_p),
                ("m", C.c_void_p), ("v", C.c_void_p), ("table_rows", C.c_int64), ("dim", C.c_int32), ("reserved", C.c_int32)]


class ExampleColumn(C.Structure):
    _fields_ = [("name", C.c_char_p), ("name_len", C.c_int32), ("kind", C.c_int32), ("n_values", C.c_int64),
                ("n_bytes", C.c_int64), ("row_counts", C.c_void_p), ("bytes_out", C.c_void_p), ("value_offsets", C.c_void_p),
                ("floats_out", C.c_void_p), ("ints_out", C.c_void_p)]


TFR_BYTES, TFR_FLOAT, TFR_INT64 = 0, 1, 2


class NativeError(RuntimeError):
    pass


_lib = None


def lib():
    """Load librf_b200.so; raises if it has not been built (python -m recommendflow_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise NativeError(f"{SO_PATH} is missing: build it with `python -m recommendflow_b200.build` "
                              "(nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(SO_PATH)
        L.rf_abi_version.restype = C.c_int
        L.rf_last_error.restype = C.c_char_p
        L.rf_launch_count.restype = C.c_int64
        L.rf_debug_fastmod.restype = C.c_uint64
        L.rf_debug_fastmod.argtypes = [C.c_uint64, C.c_uint64]
        L.rf_hash_strings.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                      C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
        L.rf_hash_strings_masked.restype = C.c_int
        L.rf_hash_strings_masked.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_char_p, C.c_int32, C.c_int,
                                             C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
        L.rf_hash_int64.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int64, C.c_int,
                                    C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
        L.rf_bag_forward.argtypes = [C.POINTER(FieldDesc), C.c_int, C.c_int64, C.c_void_p]
        L.rf_bag_forward_ex.restype = C.c_int
        L.rf_bag_forward_ex.argtypes = [C.POINTER(FieldDesc), C.c_int, C.c_int64, C.c_int, C.c_void_p]
        L.rf_release_captured_launches.restype = C.c_int
        L.rf_bag_backward.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_int64, C.c_int32,
                                      C.c_int, C.c_float, C.c_void_p, C.c_void_p]
        L.rf_sdpa_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                      C.c_void_p, C.c_void_p]
        L.rf_sdpa_forward_tc.argtypes = L.rf_sdpa_forward.argtypes
        L.rf_sdpa_forward_tc_strided.restype = C.c_int
        L.rf_sdpa_forward_tc_strided.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32,
                                                 C.c_int32, C.c_void_p, C.c_void_p]
        L.rf_dense_forward_tc.restype = C.c_int
        L.rf_dense_forward_tc.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_int,
                                          C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
        L.rf_inbatch_workspace_bytes.restype = C.c_int64
        L.rf_inbatch_workspace_bytes.argtypes = [C.c_int64]
        L.rf_inbatch_rowstats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_float,
                                          C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p]
        L.rf_inbatch_rowstats_tc.argtypes = L.rf_inbatch_rowstats.argtypes
        L.rf_inbatch_rowstats_bf16.argtypes = L.rf_inbatch_rowstats.argtypes
        L.rf_inbatch_rowstats_bf16.restype = C.c_int
        L.rf_inbatch_workspace_bytes_tc.restype = C.c_int64
        L.rf_inbatch_workspace_bytes_tc.argtypes = [C.c_int64, C.c_int32]
        L.rf_shard_route.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                     C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p]
        L.rf_shard_route_keys.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_uint64, C.c_uint64,
                                          C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                          C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p]
        L.rf_shard_route_tiles.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_uint64,
                                           C.c_uint64, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int,
                                           C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p]
        L.rf_shard_exchange_bytes.restype = C.c_int64
        L.rf_shard_exchange_bytes.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_int32]
        L.rf_sharded_bag_forward.restype = C.c_int
        L.rf_sharded_bag_forward.argtypes = [C.POINTER(ShardCtx), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                             C.c_uint64, C.c_uint64, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_int64, C.c_int,
                                             C.c_uint64, C.c_void_p, C.c_int64, C.c_void_p]
        L.rf_shard_route_tiles_ex.restype = C.c_int
        L.rf_shard_route_tiles_ex.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_uint64,
                                              C.c_uint64, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int,
                                              C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.c_void_p]
        L.rf_combine_partials.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int32, C.c_int, C.c_int32, C.c_void_p,
                                          C.c_void_p, C.c_int64, C.c_void_p]
        L.rf_bag_adam_workspace_bytes.restype = C.c_int64
        L.rf_bag_adam_workspace_bytes.argtypes = [C.c_int64, C.c_int64]
        L.rf_bag_backward_adam.restype = C.c_int
        L.rf_bag_backward_adam.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_int64, C.c_int32,
                                           C.c_int, C.POINTER(AdamParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                           C.c_void_p, C.c_int64, C.c_void_p]
        L.rf_bag_adam_multi_workspace_bytes.restype = C.c_int64
        L.rf_bag_adam_multi_workspace_bytes.argtypes = [C.POINTER(AdamField), C.c_int]
        L.rf_bag_backward_adam_multi.restype = C.c_int
        L.rf_bag_backward_adam_multi.argtypes = [C.POINTER(AdamField), C.c_int, C.c_int64, C.POINTER(AdamParams), C.c_void_p,
                                                 C.c_int64, C.c_void_p]
        L.rf_crc32c.restype = C.c_uint32
        L.rf_crc32c.argtypes = [C.c_void_p, C.c_int64]
        L.rf_masked_crc32c.restype = C.c_uint32
        L.rf_masked_crc32c.argtypes = [C.c_void_p, C.c_int64]
        L.rf_tfrecord_index.restype = C.c_int
        L.rf_tfrecord_index.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
        L.rf_example_parse_columns.restype = C.c_int
        L.rf_example_parse_columns.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(ExampleColumn), C.c_int, C.c_int]
        L.rf_sdpa_backward.restype = C.c_int
        L.rf_sdpa_backward.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.c_int32, C.c_int32] + [C.c_void_p] * 4
        L.rf_inbatch_softmax_ce_backward.restype = C.c_int
        L.rf_inbatch_softmax_ce_backward.argtypes = [C.c_void_p] * 4 + [C.c_int64, C.c_int32, C.c_float, C.c_float] + [C.c_void_p] * 3
        L.rf_inbatch_softmax_ce_backward_block.restype = C.c_int
        L.rf_inbatch_softmax_ce_backward_block.argtypes = [C.c_void_p] * 4 + [C.c_int64, C.c_int32, C.c_float, C.c_float, C.c_int] + [C.c_void_p] * 3
        L.rf_inbatch_ce_backward_tc_workspace_bytes.restype = C.c_int64
        L.rf_inbatch_ce_backward_tc_workspace_bytes.argtypes = [C.c_int64, C.c_int32]
        L.rf_inbatch_softmax_ce_backward_tc.restype = C.c_int
        L.rf_inbatch_softmax_ce_backward_tc.argtypes = [C.c_void_p] * 4 + [C.c_int64, C.c_int32, C.c_float, C.c_float, C.c_int,
                                                                          C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.rf_bag_minmax_key_grads.restype = C.c_int
        L.rf_bag_minmax_key_grads.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p,
                                              C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.rf_sdpa_backward_strided.restype = C.c_int
        L.rf_sdpa_backward_strided.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                               C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.rf_sdpa_backward_tc.restype = C.c_int
        L.rf_sdpa_backward_tc.argtypes = L.rf_sdpa_backward_strided.argtypes
        L.rf_dense_tc_workspace_bytes.restype = C.c_int64
        L.rf_dense_tc_workspace_bytes.argtypes = [C.c_int64, C.c_int32, C.c_int32]
        L.rf_dense_forward_tc_ex.restype = C.c_int
        L.rf_dense_forward_tc_ex.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_int,
                                             C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
        L.rf_tower_train_workspace_bytes.restype = C.c_int64
        L.rf_tower_train_workspace_bytes.argtypes = [C.c_int64, C.c_int32]
        L.rf_column_stats.restype = C.c_int
        L.rf_column_stats.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_int64, C.c_void_p]
        L.rf_activation_backward.restype = C.c_int
        L.rf_activation_backward.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.rf_batchnorm_backward.restype = C.c_int
        L.rf_batchnorm_backward.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                            C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.rf_vocab_build.argtypes = [C.POINTER(VocabDesc), C.c_void_p]
        L.rf_vocab_lookup_strings.argtypes = [C.POINTER(VocabDesc), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.rf_vocab_lookup_int64.argtypes = [C.POINTER(VocabDesc), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.rf_bucketize_f32.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        for name in ("rf_vocab_build", "rf_vocab_lookup_strings", "rf_vocab_lookup_int64", "rf_bucketize_f32"):
            getattr(L, name).restype = C.c_int
        for name in ("rf_hash_strings", "rf_hash_int64", "rf_bag_forward", "rf_shard_route", "rf_shard_route_keys", "rf_shard_route_tiles", "rf_set_bag_grid_limit", "rf_bag_backward", "rf_sdpa_forward", "rf_sdpa_forward_tc", "rf_inbatch_rowstats", "rf_inbatch_rowstats_tc",
                     "rf_combine_partials"):
            getattr(L, name).restype = C.c_int
        _lib = L
    return _lib


def check(rc):
    if rc == RF_OK:
        return
    msg = lib().rf_last_error().decode(errors="replace")
    if rc == RF_ERR_INVALID:
        raise ValueError(msg)
    if rc == RF_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise NativeError(msg)


def string_mask(mask_value):
    """Keras `Hashing(mask_value=...)` for string keys -> (rf_mask_mode, mask bytes)."""
    if mask_value is None:
        return MASK_NONE, b""
    raw = mask_value if isinstance(mask_value, bytes) else str(mask_value).encode("utf-8")
    if len(raw) == 0:
        return MASK_EMPTY_STRING, b""
    if len(raw) > MAX_MASK_BYTES:
        raise NotImplementedError(f"a string mask_value takes at most {MAX_MASK_BYTES} bytes, got {len(raw)}")
    return MASK_STRING_VALUE, raw


def launch_count():
    return int(lib().rf_launch_count())


def salt_to_key(salt):
    """Keras `Hashing(salt=...)`: None -> fast hash; int s -> key (s, s); [a, b] -> key (a, b)."""
    if salt is None:
        return 0, 0, 0
    if isinstance(salt, int):
        return 1, salt & (2**64 - 1), salt & (2**64 - 1)
    if isinstance(salt, (tuple, list)) and len(salt) == 2:
        return 1, int(salt[0]) & (2**64 - 1), int(salt[1]) & (2**64 - 1)
    raise ValueError(f"`salt` should be a tuple or list of two strong hash keys or a single integer. Received: salt={salt}.")


class _NoSwitch(object):
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NO_SWITCH = _NoSwitch()


def on_device(device):
    """`with on_device(dev):` = torch.cuda.device(dev) only when dev is not already the current device (the torch context
    manager costs ~10 us of host time per launch -- more than the ctypes call it wraps)."""
    import torch
    if device is None:
        return _NO_SWITCH
    if not isinstance(device, torch.device):
        device = torch.device(device)
    idx = device.index
    if idx is None or idx == torch.cuda.current_device():
        return _NO_SWITCH
    return torch.cuda.device(device)
