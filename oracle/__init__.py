"""CPU ORACLE bindings (ctypes over oracle/librf_oracle.so) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  Nothing under recommendflow_b200/ does.  See the header of
oracle/rf_oracle.c for what is restated (reference file:line) and for the parity-pinning
statement (public TF/Keras KATs pin the hashes; FarmHash >16-byte branches are unpinned).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "librf_oracle.so")
_lib = None

COMBINERS = {"sum": 0, "avg": 1, "min": 2, "max": 3}


def build(force=False):
    """Compile the C restatement (gcc).  Building the checker is not using it."""
    src = os.path.join(_HERE, "rf_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        os.environ.setdefault("OMP_WAIT_POLICY", "passive")   # idle OpenMP threads sleep instead of spinning
        _lib = C.CDLL(_SO)
        _lib.rfo_fingerprint64.restype = C.c_uint64
        _lib.rfo_fingerprint64.argtypes = [C.c_char_p, C.c_uint64]
        _lib.rfo_siphash24.restype = C.c_uint64
        _lib.rfo_siphash24.argtypes = [C.c_uint64, C.c_uint64, C.c_char_p, C.c_uint64]
        _lib.rfo_inbatch_softmax_ce.restype = C.c_double
        _lib.rfo_num_threads.restype = C.c_int
    return _lib


def _p(a, t=C.c_void_p):
    return None if a is None else a.ctypes.data_as(t)


def fingerprint64(b: bytes) -> int:
    return lib().rfo_fingerprint64(b, len(b))


def siphash24(k0: int, k1: int, b: bytes) -> int:
    return lib().rfo_siphash24(k0, k1, b, len(b))


def encode_strings(values):
    """list of str/bytes -> (uint8 arena, int32 offsets[n+1])."""
    enc = [v.encode() if isinstance(v, str) else bytes(v) for v in values]
    offs = np.zeros(len(enc) + 1, dtype=np.int32)
    if enc:
        offs[1:] = np.cumsum([len(e) for e in enc])
    arena = np.frombuffer(b"".join(enc), dtype=np.uint8).copy() if offs[-1] else np.zeros(0, np.uint8)
    return arena, offs


def _salt(salt):
    if salt is None:
        return 0, 0, 0
    if isinstance(salt, (int, np.integer)):
        return 1, int(salt), int(salt)
    return 1, int(salt[0]), int(salt[1])


def hash_strings(arena, offs, num_bins, mask_value=None, salt=None):
    """Keras Hashing over a string column.  Returns int64[n]."""
    n = len(offs) - 1
    out = np.empty(n, dtype=np.int64)
    strong, k0, k1 = _salt(salt)
    arena = np.ascontiguousarray(arena, dtype=np.uint8)
    offs = np.ascontiguousarray(offs, dtype=np.int32)
    has_mask = mask_value is not None
    m = (mask_value.encode() if isinstance(mask_value, str) else mask_value) if has_mask else b""
    lib().rfo_hash_strings(_p(arena), _p(offs), C.c_int64(n), C.c_int64(num_bins), C.c_int(strong),
                           C.c_uint64(k0), C.c_uint64(k1), C.c_int(has_mask), C.c_char_p(m),
                           C.c_int32(len(m)), _p(out))
    return out


def hash_ints(vals, num_bins, mask_value=None, salt=None):
    vals = np.ascontiguousarray(vals, dtype=np.int64).ravel()
    out = np.empty(vals.size, dtype=np.int64)
    strong, k0, k1 = _salt(salt)
    has_mask = mask_value is not None
    lib().rfo_hash_ints(_p(vals), C.c_int64(vals.size), C.c_int64(num_bins), C.c_int(strong),
                        C.c_uint64(k0), C.c_uint64(k1), C.c_int(has_mask),
                        C.c_int64(int(mask_value) if has_mask else 0), _p(out))
    return out


def bag_pool(ids, W, combiner="sum", L=None, bag_offsets=None):
    """ids flat int64; dense bags of length L or CSR bag_offsets[B+1].  Returns [B, D] fp32."""
    ids = np.ascontiguousarray(ids, dtype=np.int64).ravel()
    W = np.ascontiguousarray(W, dtype=np.float32)
    N, D = W.shape
    if bag_offsets is not None:
        bag_offsets = np.ascontiguousarray(bag_offsets, dtype=np.int32)
        B, L = len(bag_offsets) - 1, 0
    else:
        B = ids.size // L if L else 0
    out = np.empty((B, D), dtype=np.float32)
    lib().rfo_bag_pool(_p(ids), C.c_int64(B), C.c_int64(L or 0), _p(bag_offsets), _p(W), C.c_int64(N),
                       C.c_int64(D), C.c_int(COMBINERS[combiner]), _p(out), C.c_int64(D))
    return out


def gather_rows(ids, W):
    ids = np.ascontiguousarray(ids, dtype=np.int64).ravel()
    W = np.ascontiguousarray(W, dtype=np.float32)
    out = np.empty((ids.size, W.shape[1]), dtype=np.float32)
    lib().rfo_gather_rows(_p(ids), C.c_int64(ids.size), _p(W), C.c_int64(W.shape[1]), _p(out))
    return out


def hashed_bag_forward(arena, offs, B, L, tables, num_bins, salts, combiner="sum", mask_empty=True,
                       bag_offsets=None, out=None, out_col=0):
    """Fused oracle forward of one hashed field with T tables -> [B, T*D] (DoubleHashingEmbedding,
    /root/reference/backend/layers/preprocess_layers.py:94-97 when T == 2)."""
    T = len(tables)
    tables = [np.ascontiguousarray(w, dtype=np.float32) for w in tables]
    D = tables[0].shape[1]
    arena = np.ascontiguousarray(arena, dtype=np.uint8)
    offs = np.ascontiguousarray(offs, dtype=np.int32)
    if out is None:
        out = np.empty((B, T * D), dtype=np.float32)
        out_col = 0
    assert out.dtype == np.float32 and out.flags.c_contiguous
    Wp = (C.c_void_p * T)(*[w.ctypes.data for w in tables])
    nb = (C.c_int64 * T)(*[int(x) for x in num_bins])
    s = [_salt(x) for x in salts]
    us = (C.c_int * T)(*[x[0] for x in s])
    k0 = (C.c_uint64 * T)(*[x[1] for x in s])
    k1 = (C.c_uint64 * T)(*[x[2] for x in s])
    if bag_offsets is not None:
        bag_offsets = np.ascontiguousarray(bag_offsets, dtype=np.int32)
    optr = C.c_void_p(out.ctypes.data + 4 * out_col)
    lib().rfo_hashed_bag_forward(_p(arena), _p(offs), C.c_int64(B), C.c_int64(L), _p(bag_offsets), C.c_int(T),
                                 Wp, nb, us, k0, k1, C.c_int(1 if mask_empty else 0), C.c_int64(D),
                                 C.c_int(COMBINERS[combiner]), optr, C.c_int64(out.shape[1]))
    return out


def sdpa(q, k, v, mask=None):
    """q,k,v [..., S, dh] fp32; mask [..., S, 1] or [..., S] (query-row mask, see rf_oracle.c)."""
    q = np.ascontiguousarray(q, dtype=np.float32)
    k = np.ascontiguousarray(k, dtype=np.float32)
    v = np.ascontiguousarray(v, dtype=np.float32)
    S, dh = q.shape[-2:]
    NB = q.size // (S * dh)
    m = None
    if mask is not None:
        m = np.ascontiguousarray(np.broadcast_to(np.asarray(mask, dtype=np.float32).reshape(q.shape[:-1]),
                                                 q.shape[:-1]), dtype=np.float32)
    out = np.empty_like(q)
    lib().rfo_sdpa(_p(q), _p(k), _p(v), _p(m), C.c_int64(NB), C.c_int64(S), C.c_int64(dh), _p(out))
    return out


def inbatch_softmax_ce(y_true, query, doc, scale=20.0):
    """batch_neg_sample_scaled_multi_class_ce_loss; returns (loss, row_lse[B], diag[B]) in float64."""
    q = np.ascontiguousarray(query, dtype=np.float32)
    d = np.ascontiguousarray(doc, dtype=np.float32)
    y = np.ascontiguousarray(y_true, dtype=np.float32).ravel()
    B, Dt = q.shape
    lse = np.empty(B, dtype=np.float64)
    diag = np.empty(B, dtype=np.float64)
    loss = lib().rfo_inbatch_softmax_ce(_p(q), _p(d), _p(y), C.c_int64(B), C.c_int64(Dt), C.c_double(scale),
                                        _p(lse), _p(diag))
    return float(loss), lse, diag


ACTIVATIONS = {None: 0, "linear": 0, "relu": 1, "selu": 2, "tanh": 3, "sigmoid": 4, "gelu": 5}


def dense(x, kernel, bias=None, activation=None):
    """Keras Dense(units, activation): act(x @ kernel + bias); kernel [in, units] (backend/blocks/mlp.py:4-15)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    w = np.ascontiguousarray(kernel, dtype=np.float32)
    b = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
    out = np.empty((x2.shape[0], w.shape[1]), dtype=np.float32)
    lib().rfo_dense(_p(x2), C.c_int64(x2.shape[0]), C.c_int64(x2.shape[1]), _p(w), _p(b), C.c_int64(w.shape[1]),
                    C.c_int(ACTIVATIONS[activation]), _p(out))
    return out.reshape(*lead, w.shape[1])


def batchnorm_inference(x, gamma, beta, mean, var, eps):
    """Keras BatchNormalization(training=False): gamma * (x - mean) / sqrt(var + eps) + beta."""
    return ((x - mean) * (gamma / np.sqrt(var + np.float32(eps))) + beta).astype(np.float32)


def tower_mlp(x, stages, eps=1e-6, l2_normalize=True):
    """create_mlp([..], dropout, activation, BatchNormalization(eps)) at inference (mlp.py:4-15, dssm.py:25-26):
    stages = [(gamma, beta, mean, var, kernel, bias, activation), ...]; then K.l2_normalize per row."""
    h = np.ascontiguousarray(x, dtype=np.float32)
    for gamma, beta, mean, var, kernel, bias, act in stages:
        if gamma is not None:
            h = batchnorm_inference(h, gamma, beta, mean, var, eps)
        h = dense(h, kernel, bias, act)
    if l2_normalize:
        h = h / np.maximum(np.sqrt((h * h).sum(axis=1, keepdims=True)), np.float32(1e-12))
    return h.astype(np.float32)


def multi_head_attention(x, mask, wq, wk, wv, num_heads):
    """MultiHeadAttention(d_model, num_heads).call(x, x, x, mask) (attention_layers.py:137-168): Dense q/k/v with
    bias, split heads, scaled_dot_product_attention with the query-row mask, merge; no output projection."""
    B, S, d = x.shape
    q, k, v = (dense(x, w, b) for w, b in (wq, wk, wv))
    depth = d // num_heads
    split = lambda t: np.ascontiguousarray(t.reshape(B, S, num_heads, depth).transpose(0, 2, 1, 3))
    m = np.broadcast_to(np.asarray(mask, dtype=np.float32).reshape(B, 1, S, 1), (B, num_heads, S, 1))
    att = sdpa(split(q), split(k), split(v), m)
    return att.transpose(0, 2, 1, 3).reshape(B, S, d)


def num_threads():
    return int(lib().rfo_num_threads())


def set_num_threads(n):
    lib().rfo_set_num_threads(C.c_int(int(n)))


def host_threads(cap=64):
    """A FIXED thread count for the CPU arms of bench.py: the cores this process may actually use --
    min(scheduler affinity, cgroup CPU quota, cap) -- so that the reference arm's denominator does not
    depend on a timing probe (round 1: the auto-tuned count moved the baseline by +-25 % between runs)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:                                   # cgroup v2: "<quota> <period>" or "max <period>"
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        try:                               # cgroup v1
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = min(n, max(1, q // per))
        except Exception:
            pass
    return max(1, min(n, cap))


def autotune_threads(probe, candidates=None):
    """Pick the OpenMP thread count that runs `probe()` fastest (containers often grant fewer
    cores than os.cpu_count() reports).  Returns (threads, seconds)."""
    import time
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if candidates is None:
        candidates, c = [], 1
        while c < ncpu:
            candidates.append(c)
            c *= 2
        candidates.append(ncpu)
    best = None
    for th in candidates:
        set_num_threads(th)
        probe()
        t0 = time.perf_counter()
        probe()
        dt = time.perf_counter() - t0
        if best is None or dt < best[1]:
            best = (th, dt)
    set_num_threads(best[0])
    return best


def vocab_lookup(keys, vocabulary):
    """Keras StringLookup / IntegerLookup(vocabulary, output_mode="int") as the reference builds them
    (backend/layers/preprocess_layers.py:148-150): term i -> i + 1, out-of-vocabulary -> 0.
    keys: flat list of bytes/str or ints.  Pure-Python dict (small cases)."""
    norm = lambda t: t.encode() if isinstance(t, str) else t
    index = {norm(t): i + 1 for i, t in enumerate(vocabulary)}
    return np.array([index.get(norm(k), 0) for k in keys], dtype=np.int64)


def bucketize(values, boundaries):
    """Keras Discretization(bin_boundaries) (preprocess_layers.py:187): std::upper_bound with
    `value < boundary`, so id = #boundaries <= value and NaN -> len(boundaries)."""
    v = np.asarray(values, dtype=np.float32)
    b = np.asarray(boundaries, dtype=np.float32)
    out = np.searchsorted(b, v, side="right").astype(np.int64)
    out[np.isnan(v)] = b.size
    return out


def bag_backward_adam(ids, grad, W, M, V, step, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7, combiner="sum", L=None,
                      bag_offsets=None, lazy=False):
    """In-place Keras Adam step on (W, M, V) from the pooled bag's gradient; returns the touched mask."""
    ids = np.ascontiguousarray(ids, dtype=np.int64).ravel()
    grad = np.ascontiguousarray(grad, dtype=np.float32)
    B, D = grad.shape
    rows = W.shape[0]
    for a in (W, M, V):
        assert a.dtype == np.float32 and a.flags.c_contiguous and a.shape == (rows, D)
    bo = None if bag_offsets is None else np.ascontiguousarray(bag_offsets, dtype=np.int32)
    G = np.zeros((rows, D), dtype=np.float32)
    touched = np.zeros(rows, dtype=np.uint8)
    lib().rfo_bag_backward_adam(_p(ids), C.c_int64(ids.size), _p(bo), C.c_int64(L or 0), C.c_int64(B), _p(grad), C.c_int64(D),
                                C.c_int(1 if combiner == "avg" else 0), C.c_float(lr), C.c_float(beta1), C.c_float(beta2),
                                C.c_float(eps), C.c_int64(step), C.c_int(1 if lazy else 0), _p(W), _p(M), _p(V), C.c_int64(rows),
                                _p(G), _p(touched))
    return touched.astype(bool)
