"""End to end on "identical synthetic TFRecords": rows -> make_tfrecord.py's value encoding -> GZIP TFRecord
file -> parse_example densification -> every preprocessing layer of the config (ONE fused launch) -> compared
with the oracle run on the strings / ints / floats decoded from the same file."""
import os

import numpy as np
import pytest
import torch

import oracle
from recommendflow_b200.backend.layers.preprocess_layers import DiscreteEmbedding, DoubleHashingEmbedding, LookupEmbedding
from recommendflow_b200.backend.utils.preprocess_utils import get_preprocess_layers
from recommendflow_b200.config_parser import Configuration
from recommendflow_b200.data import tfrecord as tfr

pytestmark = pytest.mark.gpu


def _row(rng):
    def seq(prefix, hi, max_n, missing=0.15):
        if rng.uniform() < missing:
            return "-1"                                              # the TSV's NaN marker -> "" (make_tfrecord.py:40)
        return ",".join(f"{prefix}{int(rng.integers(0, hi))}" for _ in range(int(rng.integers(1, max_n + 1))))
    cats = ["game", "app", "book", "movie"]
    return {"clk_items": seq("i", 5000, 7), "clk_cates": seq("c", 40, 5), "uid": str(int(rng.integers(0, 10**9))),
            "item_id": seq("it", 10**6, 1), "cate_id": seq("k", 300, 1), "shop_id": seq("s", 2000, 4),
            "top_cat": ",".join(cats[int(i)] for i in rng.integers(0, 4, size=int(rng.integers(1, 3)))),
            "city_level": ",".join(str(int(v)) for v in rng.integers(0, 8, size=int(rng.integers(1, 4)))),
            "price": ",".join(f"{v:.3f}" for v in rng.uniform(0, 1500, size=int(rng.integers(1, 3)))),
            "avg_price": f"{rng.uniform(0, 200):.2f}", "dropped": "zz", "label": str(int(rng.integers(0, 2)))}


def test_tfrecord_batches_through_every_preprocess_layer(golden_dir, tmp_path):
    rng = np.random.default_rng(77)
    conf = Configuration(os.path.join(golden_dir, "configs", "synth_mixed.yaml"))
    for f in conf.features.features:
        if f.name == "uid":          # int keys + mask_value="" raises in Keras' Hashing exactly as it does here
            f.working = False
    rows = [_row(rng) for _ in range(150)]
    path = str(tmp_path / "part-0.tfr.gz")
    tfr.dump_tfrecord_data(rows, path, conf)
    layers = get_preprocess_layers(conf)
    assert "uid" not in layers and len(layers) == 9
    n_batches = 0
    for batch, labels in tfr.load_tfrecord(path, conf, batch_size=64):
        n_batches += 1
        res = layers.forward_all(batch)
        B = labels["label"].shape[0]
        for name, layer in layers.items():
            got = res[name].cpu().numpy()
            if isinstance(layer, DoubleHashingEmbedding):
                col = batch[name]
                L = col.shape[1]
                arena, offs = oracle.encode_strings(col.tolist())
                want = oracle.hashed_bag_forward(arena, offs, B, L, layer.get_weights(), [layer.num_bins] * 2, [2022, 2023],
                                                 layer.combiner)
            elif isinstance(layer, LookupEmbedding):
                if layer.key_type == "str":
                    L = batch[name].shape[1]
                    ids = oracle.vocab_lookup(batch[name].tolist(), layer.vocabulary)
                else:
                    L = batch[name].shape[1]
                    ids = oracle.vocab_lookup(batch[name].numpy().ravel().tolist(), layer.vocabulary)
                want = oracle.bag_pool(ids, layer.embedding.get_weights()[0], layer.pooling, L=L)
            else:
                assert isinstance(layer, DiscreteEmbedding)
                x = batch[name].numpy()
                x = x if x.ndim == 2 else x[:, None]
                want = oracle.bag_pool(oracle.bucketize(x.ravel(), layer.bin_boundaries), layer.embedding.get_weights()[0],
                                       layer.pooling, L=x.shape[1])
            assert got.shape == want.shape, name
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), name
    assert n_batches == 3
