"""GPU hash kernels (through the C-ABI) vs the CPU oracle: bit-exact bucket ids."""
import numpy as np
import pytest
import torch

import oracle
from recommendflow_b200.bag_ops import hash_ints, hash_strings
from recommendflow_b200.backend.layers.preprocess_layers import Hashing
from recommendflow_b200.strings import StringColumn
from tests.gpu_util import column, random_strings
from tests.test_oracle_kat import KERAS_KATS

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("values,num_bins,mask,salt,want", KERAS_KATS)
def test_public_kats_on_gpu(values, num_bins, mask, salt, want):
    layer = Hashing(num_bins, mask_value=mask, salt=salt)
    if isinstance(values[0], int):
        got = layer(torch.tensor(values, dtype=torch.int64).view(-1, 1).cuda())
    else:
        got = layer([[v] for v in values])
    assert got.view(-1).tolist() == want


@pytest.mark.parametrize("salt", [None, [2022, 2022], [0x0706050403020100, 0x0F0E0D0C0B0A0908]])
@pytest.mark.parametrize("num_bins,mask", [(1, ""), (2, ""), (3, None), (3000, ""), (100000, ""), (1000000, ""),
                                           (2**31, None), (2**32 - 1, ""), (999983, None)])
def test_random_strings_every_length_branch(salt, num_bins, mask):
    rng = np.random.default_rng(20260101)
    arena, offs = random_strings(rng, 6000, max_len=150, empty_frac=0.05)
    want = oracle.hash_strings(arena, offs, num_bins, mask, salt)
    got = hash_strings(column(arena, offs, (6000, 1)), num_bins, mask, salt).view(-1).cpu().numpy()
    assert np.array_equal(got, want)


@pytest.mark.parametrize("salt", [None, [133, 137]])
@pytest.mark.parametrize("mask", ["omar", "x", "0123456789abcdef0123456789abcdef", "ab\x00cd", "日本"])
def test_arbitrary_string_mask_value(mask, salt):
    """Keras Hashing(mask_value="<string>"): keys whose BYTES equal the mask go to bucket 0, everything else to
    1 + h mod (N - 1) -- prefixes, extensions and same-length near misses of the mask included."""
    rng = np.random.default_rng(len(mask))
    raw = mask.encode("utf-8")
    near = [raw, raw[:-1], raw + b"!", raw[:-1] + bytes([raw[-1] ^ 1]), b"", raw * 2, bytes([raw[0] ^ 0x20]) + raw[1:]]
    vals = [near[i % len(near)] if i % 3 == 0 else bytes(rng.integers(1, 256, size=rng.integers(0, 40), dtype=np.uint8)) for i in range(4000)]
    arena, offs = oracle.encode_strings(vals)
    want = oracle.hash_strings(arena, offs, 1000, raw, salt)
    assert (want == 0).sum() >= 4000 // 21
    got = hash_strings(column(arena, offs, (4000, 1)), 1000, mask, salt).view(-1).cpu().numpy()
    assert np.array_equal(got, want)
    # the staged (shared-memory) and unstaged (global-memory) byte sources agree: long filler keys force the latter
    vals2 = [v if i % 2 else v + b"#" * 600 for i, v in enumerate(vals)]
    arena2, offs2 = oracle.encode_strings(vals2)
    got2 = hash_strings(column(arena2, offs2, (4000, 1)), 1000, mask, salt).view(-1).cpu().numpy()
    assert np.array_equal(got2, oracle.hash_strings(arena2, offs2, 1000, raw, salt))
    with pytest.raises(NotImplementedError):
        hash_strings(column(arena, offs, (4000, 1)), 1000, "m" * 33, salt)


def test_double_hashing_layer_with_string_mask_value():
    from recommendflow_b200.backend.layers.preprocess_layers import DoubleHashingEmbedding
    rng = np.random.default_rng(5)
    B, L, N, D = 64, 3, 500, 8
    rows = [[("pad" if rng.uniform() < 0.3 else f"k{rng.integers(0, 1000)}") for _ in range(L)] for _ in range(B)]
    layer = DoubleHashingEmbedding(num_bins=N, output_dim=D, seeds=[2022, 2023], combiner="avg", mask_value="pad", name="h")
    ws = [rng.uniform(-0.05, 0.05, size=(N, D)).astype(np.float32) for _ in range(2)]
    layer.set_weights(ws)
    got = layer(rows).cpu().numpy()
    arena, offs = oracle.encode_strings([x for r in rows for x in r])
    want = np.concatenate([oracle.bag_pool(oracle.hash_strings(arena, offs, N, b"pad", s), w, "avg", L=L)
                           for s, w in zip([2022, 2023], ws)], axis=1)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_long_keys_take_the_unstaged_path():
    # > 16 KiB of key bytes per 1024-key round: the kernel hashes straight from global memory
    rng = np.random.default_rng(3)
    arena, offs = random_strings(rng, 3000, max_len=700)
    for salt in (None, [7, 9]):
        want = oracle.hash_strings(arena, offs, 1000003, "", salt)
        got = hash_strings(column(arena, offs, (3000, 1)), 1000003, "", salt).view(-1).cpu().numpy()
        assert np.array_equal(got, want)


def test_unaligned_arena_views():
    rng = np.random.default_rng(4)
    arena, offs = random_strings(rng, 2500, max_len=40, alphabet=b"abcdefghijklmnopqrstuvwxyz0123456789_")
    want = oracle.hash_strings(arena, offs, 77777, "", [2022, 2023])
    for lead in (1, 2, 3, 5, 9, 15):
        buf = torch.zeros(lead + arena.size + 16, dtype=torch.uint8)
        buf[lead:lead + arena.size] = torch.from_numpy(arena)
        dev = buf.cuda()
        col = StringColumn(dev[lead:], torch.from_numpy(offs).cuda(), (2500, 1))
        got = hash_strings(col, 77777, "", [2022, 2023]).view(-1).cpu().numpy()
        assert np.array_equal(got, want), lead


def test_int_keys_match_as_string_hashing():
    rng = np.random.default_rng(5)
    vals = np.concatenate([rng.integers(-2**63, 2**63 - 1, size=5000, dtype=np.int64),
                           rng.integers(-1000, 1000, size=5000, dtype=np.int64),
                           np.array([0, -1, 1, 9, 10, 99, 100, 2**63 - 1, -2**63], dtype=np.int64)])
    for salt, mask in [(None, None), ([2022, 2023], 0), (None, -1), (133, None)]:
        want = oracle.hash_ints(vals, 100000, mask, salt)
        got = hash_ints(torch.from_numpy(vals).cuda(), 100000, mask, salt).cpu().numpy()
        assert np.array_equal(got, want)


def test_realistic_feature_keys_c2_shape():
    # the C2 key shape: "fNN_<v>" ASCII keys, N = 1M bins, both hash kinds
    rng = np.random.default_rng(20260101 + 3)
    v = rng.integers(0, 10**7, size=20000)
    strs = [f"f03_{x}" for x in v]
    arena, offs = oracle.encode_strings(strs)
    for salt in (None, [2022, 2022], [2023, 2023]):
        want = oracle.hash_strings(arena, offs, 1000000, "", salt)
        got = hash_strings(column(arena, offs, (5000, 4)), 1000000, "", salt).view(-1).cpu().numpy()
        assert np.array_equal(got, want)
        assert want.min() >= 1 and want.max() <= 999999


def test_empty_input():
    col = StringColumn.from_lists([]).to("cuda")
    assert hash_strings(col, 10, "", None).numel() == 0
