"""The committed golden vectors of the hot path (tests/golden/path_golden.json, made by make_path_golden.py from
the KAT-pinned oracle): the C oracle and the independent pure-Python hashes must still reproduce them (CPU), and so
must the CUDA path through the C-ABI (GPU)."""
import json
import os

import numpy as np
import pytest

import oracle
from oracle import pyhash


@pytest.fixture(scope="module")
def golden(golden_dir):
    with open(os.path.join(golden_dir, "path_golden.json")) as fh:
        return json.load(fh)


def _tables(g):
    a = g["app_id"]
    return [np.random.default_rng(s).uniform(-0.05, 0.05, size=(a["num_bins"], a["dim"])).astype(np.float32)
            for s in a["table_rng_seeds"]]


def test_oracle_and_pyhash_reproduce_the_golden_vectors(golden):
    strings = [bytes.fromhex(h) for h in golden["strings_hex"]]
    for s, fp, s1, s2 in zip(strings, golden["fingerprint64"], golden["siphash24_2022"], golden["siphash24_2023"]):
        assert oracle.fingerprint64(s) == int(fp) == pyhash.fingerprint64(s)
        assert oracle.siphash24(2022, 2022, s) == int(s1) == pyhash.siphash24(2022, 2022, s)
        assert oracle.siphash24(2023, 2023, s) == int(s2)
    arena, offs = oracle.encode_strings(strings)
    for cfg in golden["hashing"]:
        assert oracle.hash_strings(arena, offs, cfg["num_bins"], cfg["mask_value"], cfg["salt"]).tolist() == cfg["ids"]
    ints = np.array([int(v) for v in golden["ints"]], dtype=np.int64)
    for cfg in golden["int_hashing"]:
        assert oracle.hash_ints(ints, cfg["num_bins"], cfg["mask_value"], cfg["salt"]).tolist() == cfg["ids"]
    a = golden["app_id"]
    flat = [x for r in a["rows"] for x in r]
    a2, o2 = oracle.encode_strings(flat)
    got = oracle.hashed_bag_forward(a2, o2, len(a["rows"]), 3, _tables(golden), [a["num_bins"]] * 2, a["seeds"], "sum")
    assert got.view(np.uint32).tolist() == a["pooled_sum_bits"]
    # empty strings are bucket 0 under mask_value=""; an all-pad bag pools 3 x row 0 of each table
    assert a["ids1"][6:9] == [0, 0, 0] and a["ids1"][2] == 0


@pytest.mark.gpu
def test_cuda_path_reproduces_the_golden_vectors(golden):
    import torch
    from recommendflow_b200.backend.layers.preprocess_layers import DoubleHashingEmbedding
    from recommendflow_b200.bag_ops import hash_ints, hash_strings
    from recommendflow_b200.strings import StringColumn
    strings = [bytes.fromhex(h) for h in golden["strings_hex"]]
    col = StringColumn.from_lists([[s] for s in strings]).to("cuda")
    for cfg in golden["hashing"]:
        got = hash_strings(col, cfg["num_bins"], cfg["mask_value"], cfg["salt"])
        assert got.view(-1).tolist() == cfg["ids"], cfg
    ints = torch.tensor([int(v) for v in golden["ints"]], dtype=torch.int64).view(-1, 1).cuda()
    for cfg in golden["int_hashing"]:
        assert hash_ints(ints, cfg["num_bins"], cfg["mask_value"], cfg["salt"]).view(-1).tolist() == cfg["ids"], cfg
    a = golden["app_id"]
    for combiner, key in (("sum", "pooled_sum_bits"), ("avg", "pooled_avg_bits")):
        layer = DoubleHashingEmbedding(a["num_bins"], a["dim"], a["seeds"], combiner, mask_value="", mask_zero=True,
                                       name="hashing_app_id")
        layer.set_weights(_tables(golden))
        got = layer(a["rows"]).cpu().numpy()
        assert got.view(np.uint32).tolist() == a[key], combiner
