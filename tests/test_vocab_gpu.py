"""Parity of the device vocabulary lookup / bucketisation (rf_vocab_*, rf_bucketize_f32) and of the
LookupEmbedding / DiscreteEmbedding layers built on them, against the oracle."""
import numpy as np
import pytest
import torch

import oracle
from recommendflow_b200.backend.layers.preprocess_layers import DiscreteEmbedding, LookupEmbedding
from recommendflow_b200.strings import StringColumn
from recommendflow_b200.vocab_ops import DeviceVocabulary, bucketize

pytestmark = pytest.mark.gpu


def _rand_terms(rng, n, max_len):
    terms = set()
    while len(terms) < n:
        ln = int(rng.integers(0, max_len + 1))
        terms.add(bytes(rng.integers(97, 101, size=ln, dtype=np.uint8)))       # tiny alphabet: many near-misses
    return sorted(terms, key=lambda t: (len(t), t))


@pytest.mark.parametrize("n_terms,max_len", [(0, 4), (1, 1), (7, 3), (1000, 9), (50000, 40), (300, 200)])
def test_string_vocabulary_lookup_is_exact(n_terms, max_len):
    rng = np.random.default_rng(n_terms + max_len)
    terms = _rand_terms(rng, n_terms, max_len)
    rng.shuffle(terms)
    vocab = DeviceVocabulary(terms, "cuda")
    B, L = 257, 6
    keys = []
    for _ in range(B * L):
        u = rng.uniform()
        if terms and u < 0.5:
            keys.append(terms[int(rng.integers(0, len(terms)))])
        elif terms and u < 0.7:                                                 # a term with one byte changed / added
            t = bytearray(terms[int(rng.integers(0, len(terms)))])
            if t and rng.uniform() < 0.5:
                t[int(rng.integers(0, len(t)))] ^= 1
            else:
                t.append(97)
            keys.append(bytes(t))
        else:
            keys.append(bytes(rng.integers(97, 123, size=int(rng.integers(0, max_len + 2)), dtype=np.uint8)))
    col = StringColumn.from_lists([keys[b * L:(b + 1) * L] for b in range(B)]).to("cuda")
    got = vocab.lookup(col).cpu().numpy()
    assert got.shape == (B, L)
    assert np.array_equal(got.ravel(), oracle.vocab_lookup(keys, terms))


def test_integer_vocabulary_lookup_is_exact():
    rng = np.random.default_rng(5)
    terms = np.unique(np.concatenate([rng.integers(-2**62, 2**62, size=4000), np.arange(-50, 50),
                                      [np.iinfo(np.int64).min, np.iinfo(np.int64).max, 0]])).tolist()
    rng.shuffle(terms)
    vocab = DeviceVocabulary(terms, "cuda")
    keys = np.concatenate([rng.choice(terms, size=3000), rng.integers(-200, 200, size=3000), rng.integers(-2**62, 2**62, size=500)])
    got = vocab.lookup(torch.from_numpy(keys).cuda().view(-1, 5)).cpu().numpy()
    assert np.array_equal(got.ravel(), oracle.vocab_lookup(keys.tolist(), terms))


def test_bucketize_matches_upper_bound():
    rng = np.random.default_rng(9)
    for edges in ([], [0.0], [1.0, 5.0, 10.0, 50.0], sorted(rng.normal(size=37).astype(np.float32).tolist())):
        x = np.concatenate([rng.normal(size=1000).astype(np.float32) * 20, np.asarray(edges, dtype=np.float32),
                            np.asarray([np.nan, np.inf, -np.inf, -0.0, 0.0], dtype=np.float32)])
        got = bucketize(torch.from_numpy(x).cuda(), torch.tensor(edges, dtype=torch.float32).cuda()).cpu().numpy()
        assert np.array_equal(got, oracle.bucketize(x, edges))


def test_lookup_and_discrete_layers_against_oracle():
    rng = np.random.default_rng(13)
    vocabs = ["game", "app", "music", "video"]
    layer = LookupEmbedding(16, "str", vocabs, vocab_size=len(vocabs), pooling="avg", name="lookup_top_cat")
    rows = [["app", "zzz", "video"], ["music"], ["", "game"]]
    out = layer(rows)
    w = layer.embedding.get_weights()[0]
    assert w.shape == (5, 16)
    flat = [x for r in rows for x in (r + [""] * (3 - len(r)))]
    ids = oracle.vocab_lookup(flat, vocabs)
    assert ids.tolist() == [2, 0, 4, 3, 0, 0, 0, 1, 0]
    want = oracle.bag_pool(ids, w, "avg", L=3)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32))
    assert layer.get_vocabulary() == ["[UNK]"] + vocabs

    ilayer = LookupEmbedding(8, "int", [3, 1, 2], pooling="sum", name="lookup_city_level")
    x = torch.tensor([[1, 7], [3, 2]], dtype=torch.int64)
    got = ilayer(x).cpu().numpy()
    want = oracle.bag_pool(oracle.vocab_lookup([1, 7, 3, 2], [3, 1, 2]), ilayer.embedding.get_weights()[0], "sum", L=2)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))

    edges = [1.0, 5.0, 10.0, 50.0]
    dlayer = DiscreteEmbedding(8, edges, vocab_size=len(edges), pooling="sum", name="discrete_price")
    price = rng.uniform(-5, 80, size=(64, 1)).astype(np.float32)
    price[:4, 0] = [1.0, 50.0, 49.999, np.nan]
    got = dlayer(torch.from_numpy(price)).cpu().numpy()
    want = oracle.bag_pool(oracle.bucketize(price.ravel(), edges), dlayer.embedding.get_weights()[0], "sum", L=1)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_forward_all_fuses_hashed_lookup_and_discrete_features():
    from recommendflow_b200 import _native as nat
    from recommendflow_b200.backend.layers.preprocess_layers import DoubleHashingEmbedding
    from recommendflow_b200.backend.utils.preprocess_utils import PreprocessLayers
    rng = np.random.default_rng(21)
    B = 300
    layers = PreprocessLayers()
    layers["uid"] = DoubleHashingEmbedding(5000, 16, [2022, 2023], "sum", mask_value="", mask_zero=True, name="hashing_uid")
    layers["top_cat"] = LookupEmbedding(8, "str", ["game", "app", "book"], vocab_size=3, pooling="sum", name="lookup_top_cat")
    layers["city_level"] = LookupEmbedding(4, "int", [1, 2, 3, 4, 5], vocab_size=5, pooling="max", name="lookup_city_level")
    layers["price"] = DiscreteEmbedding(8, [0.5, 10, 100.25, 1000], vocab_size=4, pooling="sum", name="discrete_price")
    cats = ["game", "app", "book", "zzz", ""]
    batch = {"uid": [[f"u{int(v)}", f"u{int(v) + 1}"] for v in rng.integers(0, 10**6, size=B)],
             "top_cat": [[cats[int(i)] for i in rng.integers(0, 5, size=3)] for _ in range(B)],
             "city_level": torch.from_numpy(rng.integers(0, 8, size=(B, 2))),
             "price": torch.from_numpy(rng.uniform(0, 2000, size=(B, 1)).astype(np.float32))}
    singles = {n: layers[n](batch[n]) for n in layers}
    before = nat.launch_count()
    res = layers.forward_all(batch)
    assert nat.launch_count() == before + 4          # 2 vocabulary lookups + 1 bucketize + ONE fused gather/pool launch
    layout, total = layers.output_layout()
    assert total == 32 + 8 + 4 + 8 and res["__fused__"].shape == (B, total)
    for n in layers:
        assert torch.equal(res[n], singles[n]), n
    # and the lookup path against the oracle
    flat = [x for r in batch["top_cat"] for x in r]
    want = oracle.bag_pool(oracle.vocab_lookup(flat, ["game", "app", "book"]), layers["top_cat"].embedding.get_weights()[0], "sum", L=3)
    assert np.array_equal(res["top_cat"].cpu().numpy().view(np.uint32), want.view(np.uint32))


def test_keras_doc_examples_on_the_device():
    # the examples of the StringLookup / IntegerLookup / Discretization docstrings (tests/test_oracle_kat.py pins the oracle on them)
    sv = DeviceVocabulary(["a", "b", "c", "d"], "cuda")
    col = StringColumn.from_lists([["a", "c", "d"], ["d", "z", "b"]]).to("cuda")
    assert sv.lookup(col).tolist() == [[1, 3, 4], [4, 0, 2]]
    iv = DeviceVocabulary([12, 36, 1138, 42], "cuda")
    assert iv.lookup(torch.tensor([[12, 1138, 42], [42, 1000, 36]], device="cuda")).tolist() == [[1, 3, 4], [4, 0, 2]]
    x = torch.tensor([[-1.5, 1.0, 3.4, .5], [0.0, 3.0, 1.3, 0.0]], device="cuda")
    assert bucketize(x, torch.tensor([0., 1., 2.], device="cuda")).tolist() == [[0, 2, 3, 1], [1, 3, 2, 1]]


def test_lookup_embedding_reference_rows_flag():
    """reference_rows=True keeps the reference's table shape (`vocab_size=len(vocabs)` rows, preprocess_utils.py factory): the
    checkpoint shape matches; the last term has no row and raises like the reference's CPU gather; the default allocates the
    extra row."""
    from recommendflow_b200.backend.layers.preprocess_layers import LookupEmbedding
    from recommendflow_b200.strings import StringColumn
    vocabs = ["a", "b", "c", "d"]
    ref = LookupEmbedding(8, "str", vocabs, vocab_size=len(vocabs), pooling="sum", name="lookup_x", reference_rows=True)
    ours = LookupEmbedding(8, "str", vocabs, vocab_size=len(vocabs), pooling="sum", name="lookup_y")
    ok = StringColumn.from_lists([["a", "c"], ["zzz", "b"]]).to("cuda")
    out = ref(ok)
    assert ref.embedding.embeddings.shape == (4, 8) and ours(ok).shape == out.shape and ours.embedding.embeddings.shape == (5, 8)
    w = ref.embedding.embeddings.detach()
    assert torch.equal(out[0], w[1] + w[3]) and torch.equal(out[1], w[0] + w[2])
    with pytest.raises(ValueError, match="is not in"):
        ref(StringColumn.from_lists([["d", "a"], ["a", "a"]]).to("cuda"))
    ours(StringColumn.from_lists([["d", "a"], ["a", "a"]]).to("cuda"))
