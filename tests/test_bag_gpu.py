"""Fused hash + gather + pool kernel (through the C-ABI and the reference-named layers) vs the
CPU oracle.  Integer results (ids) and fp32 pooled outputs are compared BIT-EXACT: the kernel
pools sequentially in index order, like the oracle (and divides `avg` with an IEEE fp32 divide).
"""
import numpy as np
import pytest
import torch

import oracle
from recommendflow_b200 import _native as nat
from recommendflow_b200.bag_ops import FieldCall, bag_forward
from recommendflow_b200.backend.layers.preprocess_layers import DoubleHashingEmbedding, EmbeddingBag
from recommendflow_b200.backend.utils.preprocess_utils import PreprocessLayers
from tests.gpu_util import column, random_strings, tables, to_dev

pytestmark = pytest.mark.gpu
ALNUM = b"abcdefghijklmnopqrstuvwxyz0123456789_"


def run_field(arena, offs, B, L, ws, num_bins, salts, combiner, mask_mode=nat.MASK_EMPTY_STRING, bag_offsets=None,
              want_ids=False):
    T, D = len(ws), ws[0].shape[1]
    col = column(arena, offs, (B, L if bag_offsets is None else None), bag_offsets)
    out = torch.full((B, T * D), float("nan"), dtype=torch.float32, device="cuda")
    ids_out = torch.empty(T, len(offs) - 1, dtype=torch.int64, device="cuda") if want_ids else None
    dev_w = to_dev(ws)
    bag_forward([FieldCall([(dev_w[t], num_bins, salts[t]) for t in range(T)], D, combiner, keys=col,
                           mask_mode=mask_mode, out=out, ids_out=ids_out, bag_len=None if bag_offsets is not None else L)], B)
    torch.cuda.synchronize()
    return out.cpu().numpy(), None if ids_out is None else ids_out.cpu().numpy()


@pytest.mark.parametrize("combiner", ["sum", "avg", "min", "max"])
@pytest.mark.parametrize("D,L,T", [(64, 4, 1), (64, 4, 2), (8, 1, 2), (16, 7, 2), (128, 50, 1), (256, 3, 2),
                                   (512, 2, 1), (12, 5, 2), (6, 3, 1), (100, 9, 2), (48, 4, 1)])
def test_dense_bags_bit_exact(combiner, D, L, T):
    rng = np.random.default_rng(1000 + D + L)
    B, N = 777, 5003
    arena, offs = random_strings(rng, B * L, max_len=20, alphabet=ALNUM, empty_frac=0.3)
    ws = tables(rng, T, N, D)
    salts = [None, [2023, 2023]][:T] if D % 8 else [[2022, 2022], [2023, 2023]][:T]
    want = oracle.hashed_bag_forward(arena, offs, B, L, ws, [N] * T, salts, combiner)
    got, ids = run_field(arena, offs, B, L, ws, N, salts, combiner, want_ids=True)
    for t in range(T):
        assert np.array_equal(ids[t], oracle.hash_strings(arena, offs, N, "", salts[t]))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("combiner", ["sum", "avg", "max"])
@pytest.mark.parametrize("D,T", [(128, 1), (64, 2), (20, 1)])
def test_jagged_bags_bit_exact(combiner, D, T):
    rng = np.random.default_rng(4242)
    B, N = 600, 100003
    lens = rng.integers(0, 201, size=B)           # includes empty bags
    lens[5] = 0
    bag = np.zeros(B + 1, dtype=np.int32)
    bag[1:] = np.cumsum(lens)
    n = int(bag[-1])
    arena, offs = random_strings(rng, n, max_len=16, alphabet=ALNUM)
    ws = tables(rng, T, N, D)
    salts = [None, [5, 6]][:T]
    want = oracle.hashed_bag_forward(arena, offs, B, 0, ws, [N] * T, salts, combiner, bag_offsets=bag)
    got, _ = run_field(arena, offs, B, 0, ws, N, salts, combiner, bag_offsets=bag)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("D", [64, 256, 10])
def test_bags_longer_than_a_round_carry_accumulators(D):
    rng = np.random.default_rng(99)
    lens = np.array([2500, 3, 1024, 1025, 0, 4096, 1], dtype=np.int64)
    bag = np.zeros(len(lens) + 1, dtype=np.int32)
    bag[1:] = np.cumsum(lens)
    arena, offs = random_strings(rng, int(bag[-1]), max_len=12, alphabet=ALNUM)
    ws = tables(rng, 2, 9973, D)
    for combiner in ("sum", "avg", "min"):
        want = oracle.hashed_bag_forward(arena, offs, len(lens), 0, ws, [9973] * 2, [[1, 2], None], combiner, bag_offsets=bag)
        got, _ = run_field(arena, offs, len(lens), 0, ws, 9973, [[1, 2], None], combiner, bag_offsets=bag)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # dense, one very long bag per sample
    B, L = 3, 1500
    arena, offs = random_strings(rng, B * L, max_len=10, alphabet=ALNUM, empty_frac=0.2)
    want = oracle.hashed_bag_forward(arena, offs, B, L, ws, [9973] * 2, [[1, 2], None], "avg")
    got, _ = run_field(arena, offs, B, L, ws, 9973, [[1, 2], None], "avg")
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_double_hashing_layer_matches_reference_semantics():
    # base_conf.yaml's one working hashed feature: app_id, N=3000, D=16, sum, seeds [2022, 2023]
    rng = np.random.default_rng(11)
    rows = [["com.app.%d" % rng.integers(0, 500) for _ in range(rng.integers(0, 4))] for _ in range(257)]
    layer = DoubleHashingEmbedding(num_bins=3000, output_dim=16, seeds=[2022, 2023], mask_value="", mask_zero=True,
                                   combiner="sum", name="hashing_app_id")
    ws = tables(rng, 2, 3000, 16)
    layer.set_weights(ws)
    got = layer(rows)
    L = max(len(r) for r in rows)
    flat = [x for r in rows for x in (r + [""] * (L - len(r)))]
    arena, offs = oracle.encode_strings(flat)
    want = oracle.hashed_bag_forward(arena, offs, len(rows), L, ws, [3000, 3000], [2022, 2023], "sum")
    assert got.shape == (257, 32)
    assert np.array_equal(got.cpu().numpy().view(np.uint32), want.view(np.uint32))
    # pads are pooled in as row 0 (reference quirk): an all-pad sample equals L * W[0]
    empty = [i for i, r in enumerate(rows) if not r][0]
    acc = np.zeros(16, np.float32)
    for _ in range(L):
        acc = acc + ws[0][0]
    assert np.array_equal(got[empty, :16].cpu().numpy(), acc)
    for w_got, w_set in zip(layer.get_weights(), ws):
        assert np.array_equal(w_got, w_set)


def test_null_first_last_combiners():
    rng = np.random.default_rng(12)
    B, L, D, N = 9, 5, 8, 101
    arena, offs = random_strings(rng, B * L, max_len=6, alphabet=ALNUM, empty_frac=0.2)
    ws = tables(rng, 2, N, D)
    ids = [oracle.hash_strings(arena, offs, N, "", s).reshape(B, L) for s in (2022, 2023)]
    E = [w[i] for w, i in zip(ws, ids)]                                  # [B, L, D] each
    col = column(arena, offs, (B, L))
    for combiner, want in [("null", np.concatenate(E, axis=1)), ("first", np.concatenate([E[0][0], E[1][0]], axis=1)),
                           ("last", np.concatenate([E[0][-1], E[1][-1]], axis=1))]:
        layer = DoubleHashingEmbedding(N, D, [2022, 2023], combiner, mask_value="", name="h")
        layer.set_weights(ws)
        got = layer(col).cpu().numpy()
        assert got.shape == want.shape and np.array_equal(got, want), combiner


def test_embedding_bag_on_ids_and_int_keys():
    rng = np.random.default_rng(13)
    B, L, D, N = 300, 6, 32, 1000
    (w,) = tables(rng, 1, N, D)
    ids = rng.integers(0, N, size=(B, L), dtype=np.int64)
    for combiner in ("sum", "avg", "min", "max"):
        bag = EmbeddingBag(N, D, combiner=combiner, name="bag")
        bag.set_weights([w])
        got = bag(torch.from_numpy(ids).cuda()).cpu().numpy()
        assert np.array_equal(got.view(np.uint32), oracle.bag_pool(ids, w, combiner, L=L).view(np.uint32))
    # integer feature through DoubleHashingEmbedding: values hashed as decimal strings, mask_value=0
    vals = rng.integers(0, 50, size=(B, L), dtype=np.int64)
    layer = DoubleHashingEmbedding(N, D, [2022, 2023], "sum", mask_value=0, name="hi")
    ws = tables(rng, 2, N, D)
    layer.set_weights(ws)
    got = layer(torch.from_numpy(vals).cuda()).cpu().numpy()
    want = np.concatenate([oracle.bag_pool(oracle.hash_ints(vals, N, 0, s), w_, "sum", L=L) for s, w_ in zip((2022, 2023), ws)], axis=1)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_fused_multi_field_launch_equals_per_layer_calls():
    rng = np.random.default_rng(14)
    B = 1500
    specs = {"f_a": (50000, 64, 4, "sum"), "f_b": (3000, 16, 1, "avg"), "f_c": (100000, 8, 9, "max"),
             "f_d": (777, 128, 2, "sum"), "f_e": (1000000, 64, 4, "min")}
    layers, batch, want = PreprocessLayers(), {}, {}
    for name, (N, D, L, comb) in specs.items():
        layer = DoubleHashingEmbedding(N, D, [2022, 2023], comb, mask_value="", mask_zero=True, name=f"hashing_{name}")
        ws = tables(rng, 2, N, D)
        layer.set_weights(ws)
        arena, offs = random_strings(rng, B * L, max_len=14, alphabet=ALNUM, empty_frac=0.25)
        layers[name] = layer
        batch[name] = column(arena, offs, (B, L))
        want[name] = oracle.hashed_bag_forward(arena, offs, B, L, ws, [N, N], [2022, 2023], comb)
    before = nat.launch_count()
    res = layers.forward_all(batch)
    assert nat.launch_count() == before + 1                      # ONE launch for all five fields
    layout, total = layers.output_layout()
    assert res["__fused__"].shape == (B, total)
    for name in specs:
        assert np.array_equal(res[name].cpu().numpy().view(np.uint32), want[name].view(np.uint32)), name
        single = layers[name](batch[name])
        assert torch.equal(single, res[name])


def test_many_fields_exceeding_the_smem_prefix_table():
    rng = np.random.default_rng(15)
    B, F = 64, 600
    calls, wants, outs = [], [], []
    ws = tables(rng, 1, 997, 8)
    dev_w = to_dev(ws)
    for f in range(F):
        arena, offs = random_strings(rng, B, max_len=8, alphabet=ALNUM)
        out = torch.empty(B, 8, device="cuda")
        calls.append(FieldCall([(dev_w[0], 997, [f, f + 1])], 8, "sum", keys=column(arena, offs, (B, 1)),
                               mask_mode=nat.MASK_EMPTY_STRING, out=out, bag_len=1))
        wants.append(oracle.hashed_bag_forward(arena, offs, B, 1, ws, [997], [[f, f + 1]], "sum"))
        outs.append(out)
    bag_forward(calls, B)
    for out, want in zip(outs, wants):
        assert np.array_equal(out.cpu().numpy(), want)


def test_large_batch_linearity_property():
    # size-independent property at C2's full batch: pooling with sum is additive in the tables,
    # sum(W1 + W2)[ids] == sum(W1)[ids] + sum(W2)[ids] when W2 = 0, and ids are in range.
    rng = np.random.default_rng(16)
    B, L, N, D = 65536, 4, 1000000, 64
    v = rng.integers(0, 10**7, size=B * L)
    digits = np.char.mod("f07_%d", v)
    arena, offs = oracle.encode_strings(digits.tolist())
    col = column(arena, offs, (B, L))
    w = torch.empty(N, D, device="cuda").uniform_(-0.05, 0.05)
    out = torch.empty(B, D, device="cuda")
    ids_out = torch.empty(1, B * L, dtype=torch.int64, device="cuda")
    bag_forward([FieldCall([(w, N, None)], D, "sum", keys=col, mask_mode=nat.MASK_EMPTY_STRING, out=out, ids_out=ids_out,
                           bag_len=L)], B)
    ids = ids_out.view(B, L)
    assert int(ids.min()) >= 1 and int(ids.max()) <= N - 1
    ref = w[ids[:, 0]]
    for l in range(1, L):
        ref = ref + w[ids[:, l]]                                   # same left-to-right order
    assert torch.equal(out, ref)
    sample = rng.integers(0, B * L, size=4096)
    want = oracle.hash_strings(arena, offs, N, "", None)[sample]
    assert np.array_equal(ids_out.view(-1).cpu().numpy()[sample], want)


def test_cuda_graph_capture_and_replay_with_new_keys():
    # A captured launch owns its descriptor slot; replaying after the key BUFFERS were refilled must
    # use the new keys (descriptors hold pointers, not data).
    rng = np.random.default_rng(17)
    B, L, N, D = 4096, 4, 50021, 64
    ws = tables(rng, 2, N, D)
    dev_w = to_dev(ws)
    arenas = [random_strings(rng, B * L, max_len=10, alphabet=ALNUM) for _ in range(2)]
    cap_bytes = max(a.size for a, _ in arenas) + 16
    data = torch.zeros(cap_bytes, dtype=torch.uint8, device="cuda")
    offs = torch.zeros(B * L + 1, dtype=torch.int32, device="cuda")
    from recommendflow_b200.strings import StringColumn
    col = StringColumn(data, offs, (B, L))
    out = torch.empty(B, 2 * D, device="cuda")
    call = [FieldCall([(dev_w[0], N, [2022, 2022]), (dev_w[1], N, [2023, 2023])], D, "sum", keys=col,
                      mask_mode=nat.MASK_EMPTY_STRING, out=out, bag_len=L)]

    def load(i):
        a, o = arenas[i]
        data[:a.size].copy_(torch.from_numpy(a))
        offs.copy_(torch.from_numpy(o))

    load(0)
    bag_forward(call, B)                      # warm-up outside capture (allocates the slot pool)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        bag_forward(call, B)
    for i in (1, 0, 1):
        load(i)
        g.replay()
        torch.cuda.synchronize()
        a, o = arenas[i]
        want = oracle.hashed_bag_forward(a, o, B, L, ws, [N, N], [[2022, 2022], [2023, 2023]], "sum")
        assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32)), i


@pytest.mark.parametrize("combiner,D", [("sum", 64), ("avg", 16), ("sum", 6)])
def test_backward_sgd_matches_numpy(combiner, D):
    # gradient of the pooled bag w.r.t. every gathered row (pads included), fused with W -= lr * g.
    # Atomics make the add order free: compare within fp32 re-association of the duplicates.  The pad row
    # collects ~1000 addends of magnitude <= 0.5, so the order-dependent error is bounded by about
    # n_dup * 2^-24 * max|partial sum| ~ 1000 * 6e-8 * 5 = 3e-4 (typically 1e-5): atol 5e-4.
    from recommendflow_b200.bag_ops import bag_backward
    rng = np.random.default_rng(18)
    B, L, N = 700, 5, 997
    arena, offs = random_strings(rng, B * L, max_len=6, alphabet=ALNUM, empty_frac=0.3)
    (w,) = tables(rng, 1, N, D)
    ids = oracle.hash_strings(arena, offs, N, "", [2022, 2022])
    g = rng.standard_normal((B, D)).astype(np.float32)
    lr = 0.1
    want = w.astype(np.float64).copy()
    np.add.at(want, ids, -lr * np.repeat(g.astype(np.float64), L, axis=0) / (L if combiner == "avg" else 1))
    col = column(arena, offs, (B, L))
    dev_w = torch.from_numpy(w).cuda()
    out = torch.empty(B, D, device="cuda")
    ids_out = torch.empty(1, B * L, dtype=torch.int64, device="cuda")
    bag_forward([FieldCall([(dev_w, N, [2022, 2022])], D, combiner, keys=col, mask_mode=nat.MASK_EMPTY_STRING, out=out,
                           ids_out=ids_out, bag_len=L)], B)
    assert np.array_equal(ids_out.cpu().numpy().ravel(), ids)
    bag_backward(ids_out[0], dev_w, torch.from_numpy(g).cuda(), -lr, combiner, bag_len=L)
    np.testing.assert_allclose(dev_w.cpu().numpy(), want, rtol=1e-5, atol=5e-4)
    # jagged bags
    lens = rng.integers(0, 9, size=B)
    bag = np.zeros(B + 1, dtype=np.int32)
    bag[1:] = np.cumsum(lens)
    n = int(bag[-1])
    jid = rng.integers(0, N, size=n)
    want = w.astype(np.float64).copy()
    scale = np.repeat(1.0 / np.maximum(lens, 1), lens) if combiner == "avg" else np.ones(n)
    np.add.at(want, jid, 0.5 * np.repeat(g.astype(np.float64), lens, axis=0) * scale[:, None])
    dev_w = torch.from_numpy(w).cuda()
    bag_backward(torch.from_numpy(jid).cuda(), dev_w, torch.from_numpy(g).cuda(), 0.5, combiner,
                 bag_offsets=torch.from_numpy(bag).cuda())
    np.testing.assert_allclose(dev_w.cpu().numpy(), want, rtol=1e-5, atol=5e-4)


@pytest.mark.parametrize("combiner", ["max", "min"])
def test_minmax_pooling_backward_matches_tensorflow_gradient_rule(combiner):
    """Backward of "min" / "max" pooling (tf.reduce_min / tf.reduce_max over the bag, preprocess_layers.py:43-68): TensorFlow's
    _MinOrMaxGrad gives the pooled element's gradient to the keys whose row element equals it, split equally among ties.
    Checked against torch autograd of amax / amin (the same tie rule) in float64, dense and jagged bags with duplicate keys,
    then applied as an SGD step through the one-key-per-bag form of rf_bag_backward."""
    from recommendflow_b200.bag_ops import bag_backward, bag_minmax_key_grads
    rng = np.random.default_rng(23)
    B, L, N, D = 300, 6, 211, 16
    (w,) = tables(rng, 1, N, D)
    w = np.round(w * 40).astype(np.float32) / 40                       # coarse values: plenty of exact ties between rows
    ids = rng.integers(0, N, size=(B, L))
    ids[:, 1] = ids[:, 0]                                              # and every bag holds a duplicated key
    g = rng.standard_normal((B, D)).astype(np.float32)
    dev_w, dev_ids, dev_g = torch.from_numpy(w).cuda(), torch.from_numpy(ids).cuda(), torch.from_numpy(g).cuda()
    out = torch.empty(B, D, device="cuda")
    bag_forward([FieldCall([(dev_w, N, None)], D, combiner, ids=dev_ids.view(1, -1), out=out, bag_len=L)], B)
    w64 = torch.from_numpy(w).double().requires_grad_(True)
    rows = w64[torch.from_numpy(ids)]                                  # [B, L, D]
    y = rows.amax(dim=1) if combiner == "max" else rows.amin(dim=1)
    assert np.array_equal(out.cpu().numpy(), y.detach().float().numpy())
    rows.retain_grad()
    (y * torch.from_numpy(g).double()).sum().backward()
    kg = bag_minmax_key_grads(dev_ids.view(-1), dev_w, out, dev_g, bag_len=L)
    np.testing.assert_allclose(kg.cpu().numpy().reshape(B, L, D), rows.grad.numpy(), rtol=1e-6, atol=1e-7)
    # the row update: W -= lr * dW, dW = scatter-add of the per-key rows (what autograd accumulated into w64.grad)
    before = dev_w.clone()
    bag_backward(dev_ids.view(-1), dev_w, kg, -0.1, "sum", bag_len=1)
    np.testing.assert_allclose(dev_w.cpu().numpy(), (before.double().cpu() - 0.1 * w64.grad).float().numpy(), rtol=1e-5, atol=1e-5)
    # jagged bags (empty bags produce nothing)
    lens = rng.integers(0, 7, size=B)
    bag = np.zeros(B + 1, dtype=np.int32)
    bag[1:] = np.cumsum(lens)
    jid = rng.integers(0, N, size=int(bag[-1]))
    dev_w = torch.from_numpy(w).cuda()
    jout = torch.empty(B, D, device="cuda")
    dev_bag = torch.from_numpy(bag).cuda()
    bag_forward([FieldCall([(dev_w, N, None)], D, combiner, ids=torch.from_numpy(jid).cuda().view(1, -1), out=jout, bag_offsets=dev_bag,
                           n_items=len(jid))], B)
    kg = bag_minmax_key_grads(torch.from_numpy(jid).cuda(), dev_w, jout, dev_g, bag_offsets=dev_bag).cpu().numpy()
    for b in rng.choice(B, size=40, replace=False):
        k0, k1 = bag[b], bag[b + 1]
        if k1 == k0:
            continue
        r = w[jid[k0:k1]].astype(np.float64)
        yb = r.max(axis=0) if combiner == "max" else r.min(axis=0)
        hit = r == yb
        np.testing.assert_allclose(kg[k0:k1], hit * (g[b] / hit.sum(axis=0)), rtol=1e-6, atol=1e-7)


def test_gapped_bags_with_explicit_ends():
    # bag_ends: bags in order but with unused gaps between them (the sharded "tile" routing layout);
    # the gaps hold garbage ids that must never be dereferenced
    rng = np.random.default_rng(19)
    B, N, D = 900, 5003, 128
    lens = rng.integers(0, 40, size=B)
    gaps = rng.integers(0, 3000, size=B) * (rng.uniform(size=B) < 0.1)         # a big gap now and then
    begin = np.zeros(B, dtype=np.int64)
    pos = 0
    for b in range(B):
        pos += gaps[b]
        begin[b] = pos
        pos += lens[b]
    ids = np.full(pos + 5, 2**40, dtype=np.int64)                               # garbage everywhere ...
    real = rng.integers(0, N, size=int(lens.sum()))
    off = 0
    for b in range(B):
        ids[begin[b]:begin[b] + lens[b]] = real[off:off + lens[b]]              # ... except inside the bags
        off += lens[b]
    (w,) = tables(rng, 1, N, D)
    csr = np.zeros(B + 1, dtype=np.int32)
    csr[1:] = np.cumsum(lens)
    for combiner in ("sum", "avg", "max"):
        want = oracle.bag_pool(real, w, combiner, bag_offsets=csr)
        out = torch.full((B, D), float("nan"), device="cuda")
        bag_forward([FieldCall([(to_dev([w])[0], N, None)], D, combiner, ids=torch.from_numpy(ids).cuda().view(1, -1),
                               bag_offsets=torch.from_numpy(begin.astype(np.int32)).cuda(),
                               bag_ends=torch.from_numpy((begin + lens).astype(np.int32)).cuda(), out=out, n_items=int(lens.sum()))], B)
        assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32)), combiner


@pytest.mark.parametrize("seed,B,max_len", [(23, 3000, 12), (29, 700, 5000), (31, 1, 7000)])
def test_gapped_bags_long_and_empty(seed, B, max_len):
    # gapped bags longer than one round (carry across rounds), runs of empty bags, several tiles
    rng = np.random.default_rng(seed)
    N, D = 7919, 64
    lens = rng.integers(0, max_len + 1, size=B) * (rng.uniform(size=B) < 0.7)
    if B > 1:
        lens[rng.integers(0, B)] = max_len
    gaps = rng.integers(0, 500, size=B) * (rng.uniform(size=B) < 0.3)
    begin = np.cumsum(gaps + np.concatenate([[0], lens[:-1]])).astype(np.int64)
    total = int(begin[-1] + lens[-1])
    ids = np.full(total + 5, 2**40, dtype=np.int64)
    real = rng.integers(0, N, size=int(lens.sum()))
    csr = np.zeros(B + 1, dtype=np.int32)
    csr[1:] = np.cumsum(lens)
    for b in range(B):
        ids[begin[b]:begin[b] + lens[b]] = real[csr[b]:csr[b + 1]]
    (w,) = tables(rng, 1, N, D)
    for combiner in ("sum", "avg"):
        want = oracle.bag_pool(real, w, combiner, bag_offsets=csr)
        out = torch.full((B, D), float("nan"), device="cuda")
        bag_forward([FieldCall([(to_dev([w])[0], N, None)], D, combiner, ids=torch.from_numpy(ids).cuda().view(1, -1),
                               bag_offsets=torch.from_numpy(begin.astype(np.int32)).cuda(),
                               bag_ends=torch.from_numpy((begin + lens).astype(np.int32)).cuda(), out=out,
                               n_items=int(lens.sum()))], B)
        assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32)), combiner


@pytest.mark.parametrize("combiner,jagged,lazy", [("sum", False, False), ("avg", False, False), ("avg", True, False),
                                                  ("sum", True, True)])
def test_backward_adam_matches_keras_semantics(combiner, jagged, lazy):
    # tf.keras Adam on the Embedding variable: duplicates summed, every row decays and moves (lazy=False).
    # Two consecutive steps so that non-zero moments are exercised.  Rows whose run is summed in key order
    # (<= 128 duplicates) and untouched rows are bit-exact; the pad-like hot rows (longer runs, CTA-reduced in
    # a different order) are compared within fp32 re-association tolerance.
    from recommendflow_b200.bag_ops import BagAdam
    rng = np.random.default_rng(37)
    N, D, B, L = 4099, 32, 600, 7
    (w,) = tables(rng, 1, N, D)
    m, v = np.zeros_like(w), np.zeros_like(w)
    dev_w = torch.from_numpy(w.copy()).cuda()
    opt = BagAdam(dev_w, learning_rate=1e-2, lazy=lazy)
    for step in (1, 2):
        if jagged:
            lens = rng.integers(0, 2 * L, size=B)
            bag = np.zeros(B + 1, dtype=np.int32)
            bag[1:] = np.cumsum(lens)
            n = int(bag[-1])
        else:
            bag, n = None, B * L
        ids = rng.integers(1, N, size=n)
        ids[rng.uniform(size=n) < 0.3] = 0                                         # the pad row collects a long run
        ids[rng.uniform(size=n) < 0.1] = 17                                        # a second hot row
        g = rng.normal(size=(B, D)).astype(np.float32)
        touched = oracle.bag_backward_adam(ids, g, w, m, v, step, lr=1e-2, combiner=combiner, L=None if jagged else L,
                                           bag_offsets=bag, lazy=lazy)
        opt.apply(torch.from_numpy(ids).cuda(), torch.from_numpy(g).cuda(), combiner, bag_len=None if jagged else L,
                  bag_offsets=None if bag is None else torch.from_numpy(bag).cuda())
        counts = np.bincount(ids, minlength=N)
        exact = counts <= 128
        assert touched.sum() == (counts > 0).sum() and (~exact).sum() >= 2
        for name, got, want in (("w", dev_w, w), ("m", opt.m, m), ("v", opt.v, v)):
            got = got.cpu().numpy()
            assert np.array_equal(got[exact].view(np.uint32), want[exact].view(np.uint32)), (name, step)
            np.testing.assert_allclose(got[~exact], want[~exact], rtol=2e-4, atol=1e-6, err_msg=f"{name} step {step}")
            want[~exact] = got[~exact]                                             # carry the device's rounding forward


@pytest.mark.parametrize("lazy", [False, True])
def test_backward_adam_multi_table_single_pass(lazy):
    # several tables of one width updated in ONE pass (global row / key numbering): each must come out exactly
    # as the single-table oracle says, including a table that received no gradient this step (Keras: it still
    # decays; lazy: untouched)
    from recommendflow_b200.bag_ops import BagAdamGroup
    rng = np.random.default_rng(41)
    D, B = 16, 300
    sizes = [1009, 17, 5003, 64]
    host = [tables(rng, 1, n, D)[0] for n in sizes]
    ms, vs = [np.zeros_like(w) for w in host], [np.zeros_like(w) for w in host]
    devs = [torch.from_numpy(w.copy()).cuda() for w in host]
    group = BagAdamGroup(devs, learning_rate=3e-3, lazy=lazy)
    before = nat.launch_count()
    for step in (1, 2, 3):
        g_all = rng.normal(size=(B, 4 * D)).astype(np.float32)              # one fused gradient buffer, 4 column slices
        dev_g = torch.from_numpy(g_all).cuda()
        updates, exact = [], []
        for i, n in enumerate(sizes):
            if i == 3 and step != 2:                                          # table 3 is only touched at step 2
                updates.append(None)
                oracle.bag_backward_adam(np.zeros(0, dtype=np.int64), np.zeros((B, D), dtype=np.float32), host[i], ms[i], vs[i],
                                         step, lr=3e-3, L=1, lazy=lazy)
                exact.append(np.ones(n, dtype=bool))
                continue
            if i % 2 == 0:                                                    # dense bags
                L, bag = 3, None
                ids = rng.integers(0, n, size=B * L)
            else:                                                             # jagged bags
                lens = rng.integers(0, 6, size=B)
                bag = np.zeros(B + 1, dtype=np.int32)
                bag[1:] = np.cumsum(lens)
                L, ids = None, rng.integers(0, n, size=int(bag[-1]))
            comb = "avg" if i == 1 else "sum"
            g = np.ascontiguousarray(g_all[:, i * D:(i + 1) * D])
            oracle.bag_backward_adam(ids, g, host[i], ms[i], vs[i], step, lr=3e-3, combiner=comb, L=L, bag_offsets=bag, lazy=lazy)
            updates.append((torch.from_numpy(ids).cuda(), dev_g[:, i * D:(i + 1) * D], comb, L,
                            None if bag is None else torch.from_numpy(bag).cuda()))
            exact.append(np.bincount(ids, minlength=n) <= 128)
        group.apply(updates, B)
        for i in range(len(sizes)):
            for name, got, want in (("w", devs[i], host[i]), ("m", group.m[i], ms[i]), ("v", group.v[i], vs[i])):
                got = got.cpu().numpy()
                ok = exact[i]
                assert np.array_equal(got[ok].view(np.uint32), want[ok].view(np.uint32)), (name, i, step)
                np.testing.assert_allclose(got[~ok], want[~ok], rtol=2e-4, atol=1e-6, err_msg=f"{name} table {i} step {step}")
                want[~ok] = got[~ok]
    assert nat.launch_count() - before == 3 * (3 if lazy else 4)              # per step: prep, rows, heavy (+ dense)


def test_apply_fused_keras_adam_with_live_row_bitmap_is_bit_exact():
    """BagAdamGroup.apply_fused (the training loop's entry: one fused gradient buffer, cached descriptors) keeps a bitmap of
    the rows that ever received a gradient, and the all-rows decay of Keras' non-lazy Adam skips the others without reading
    them.  Every table, both moments: bit for bit what the oracle's full pass gives, over steps that touch new rows."""
    from recommendflow_b200.bag_ops import BagAdamGroup
    rng = np.random.default_rng(77)
    D, B = 8, 256
    sizes, lens, combs = [5003, 1009, 40000], [3, 1, 2], ["sum", "avg", "sum"]
    host = [tables(rng, 1, n, D)[0] for n in sizes]
    ms, vs = [np.zeros_like(w) for w in host], [np.zeros_like(w) for w in host]
    devs = [torch.from_numpy(w.copy()).cuda() for w in host]
    group = BagAdamGroup(devs, learning_rate=2e-3)
    ids_dev = [torch.zeros(B * L, dtype=torch.int64, device="cuda") for L in lens]       # static id buffers, refilled per step
    for step in range(1, 6):
        g_all = rng.normal(size=(B, 3 * D)).astype(np.float32)
        dev_g = torch.from_numpy(g_all).cuda()
        for i, (n, L) in enumerate(zip(sizes, lens)):
            ids = rng.integers(0, min(n, 300 * step), size=B * L)                          # the touched set grows step by step
            ids_dev[i].copy_(torch.from_numpy(ids))
            oracle.bag_backward_adam(ids, np.ascontiguousarray(g_all[:, i * D:(i + 1) * D]), host[i], ms[i], vs[i], step, lr=2e-3,
                                     combiner=combs[i], L=L)
        group.apply_fused(ids_dev, dev_g, [0, D, 2 * D], combs, lens, B)
        live = np.unpackbits(group._live.cpu().numpy().view(np.uint8), bitorder="little")
        base = 0
        for i, n in enumerate(sizes):
            touched = (np.abs(ms[i]).sum(axis=1) > 0) | (np.abs(vs[i]).sum(axis=1) > 0)
            assert np.array_equal(live[base:base + n].astype(bool) | ~touched, np.ones(n, dtype=bool)), "every row with a moment is marked"
            assert live[base:base + n].sum() < n, "and cold rows stay unmarked"
            base += n
            for name, got, want in (("w", devs[i], host[i]), ("m", group.m[i], ms[i]), ("v", group.v[i], vs[i])):
                assert np.array_equal(got.cpu().numpy().view(np.uint32), want.view(np.uint32)), (name, i, step)
