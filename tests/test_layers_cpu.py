"""Host-side behaviour of the reference-named layers that needs no GPU: names, configs, errors."""
import os

import numpy as np
import pytest
import torch

from recommendflow_b200 import _native as nat
from recommendflow_b200.backend.layers.preprocess_layers import (DiscreteEmbedding, DoubleHashingEmbedding, EmbeddingBag,
                                                                 Hashing, LookupEmbedding)
from recommendflow_b200.backend.utils.preprocess_utils import get_preprocess_layers
from recommendflow_b200.config_parser import Configuration
from recommendflow_b200.strings import ARENA_SLACK, StringColumn


def test_double_hashing_names_and_config():
    layer = DoubleHashingEmbedding(num_bins=3000, output_dim=16, seeds=[2022, 2023], combiner="sum", mask_value="",
                                   mask_zero=True, name="hashing_app_id")
    assert layer.name == "hashing_app_id"
    assert (layer.hash1.name, layer.hash2.name) == ("hashing_app_id_hashing1", "hashing_app_id_hashing2")
    assert (layer.emb1.name, layer.emb2.name) == ("hashing_app_id_embedding_bag1", "hashing_app_id_embedding_bag2")
    assert (layer.hash1.salt, layer.hash2.salt) == (2022, 2023)
    cfg = layer.get_config()
    assert cfg["combiner"] == "sum" and cfg["seeds"] == [2022, 2023]
    assert layer.emb1.get_config()["combiner"] == "sum"


def test_constructor_errors_match_reference():
    with pytest.raises(ValueError, match="`num_bins` cannot be `None` or non-positive values."):
        DoubleHashingEmbedding(0, 8, [1, 2], "sum")
    with pytest.raises(ValueError, match="`num_bins` cannot be `None` or non-positive values."):
        DoubleHashingEmbedding(None, 8, [1, 2], "sum")
    with pytest.raises(TypeError):        # the reference indexes the raw int seed (preprocess_layers.py:89)
        DoubleHashingEmbedding(10, 8, 2022, "sum")
    with pytest.raises(ValueError):
        Hashing(0)
    with pytest.raises(ValueError, match="Unsupported type for lookup feature"):
        LookupEmbedding(8, "float", [1.0], name="lookup_x")


def test_unknown_combiner_raises_reference_message():
    bag = EmbeddingBag(10, 4, combiner="median", name="b")
    with pytest.raises(ValueError, match="Do not support combiner = 'median', supported: \\[null, sum, min, max, avg, first, last\\]"):
        bag._check_combiner()


def test_factory_builds_reference_layer_set(golden_dir):
    conf = Configuration(os.path.join(golden_dir, "configs", "synth_mixed.yaml"))
    layers = get_preprocess_layers(conf)
    assert list(layers) == ["clk_items", "clk_cates", "uid", "item_id", "cate_id", "shop_id", "top_cat", "city_level",
                            "price", "avg_price"]
    uid = layers["uid"]
    assert isinstance(uid, DoubleHashingEmbedding) and uid.name == "hashing_uid"
    assert (uid.num_bins, uid.output_dim, uid.mask_value, uid.combiner, uid.seeds) == (1000000, 32, "", "sum", [2022, 2023])
    assert isinstance(layers["top_cat"], LookupEmbedding) and layers["top_cat"].name == "lookup_top_cat"
    assert isinstance(layers["price"], DiscreteEmbedding) and layers["price"].name == "discrete_price"
    assert layers["price"].embedding.name == "discrete_price_disc_lookup_embedding"
    layout, total = layers.output_layout()
    assert layout["clk_items"] == (0, 128) and layout["uid"] == (256, 64)
    # pooled lookup / discrete features join the fused buffer after the hashed ones
    assert layout["top_cat"] == (416, 8) and layout["city_level"] == (424, 4) and layout["avg_price"] == (436, 8) and total == 444
    # the reference's factory passes vocab_size=len(vocabs) although ids go up to len(vocabs); the table here
    # gets the one extra row so that the last term stays inside it
    assert layers["top_cat"].embedding.input_dim == len(layers["top_cat"].vocabulary) + 1 == 4
    with pytest.raises(ValueError, match="repeated term"):
        from recommendflow_b200.vocab_ops import DeviceVocabulary
        DeviceVocabulary(["a", "b", "a"], "cuda")


def test_string_column_padding_and_slack():
    col = StringColumn.from_lists([["a", "bc"], ["def"], []])
    assert col.shape == (3, 2) and col.n_items == 6
    assert col.tolist() == [b"a", b"bc", b"def", b"", b"", b""]
    assert col.data.numel() == 6 + ARENA_SLACK and col.offsets.dtype == torch.int32
    jag = StringColumn.from_lists([["a", "bc"], ["def"], []], jagged=True)
    assert jag.shape == (3, None) and jag.bag_offsets.tolist() == [0, 2, 3, 3] and jag.n_items == 3
    arr = StringColumn.from_numpy(np.array([["x", ""], ["yy", "z"]], dtype=object))
    assert arr.tolist() == [b"x", b"", b"yy", b"z"]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    layer = DoubleHashingEmbedding(100, 8, [1, 2], "sum", mask_value="", name="h")
    with pytest.raises(nat.NativeError, match="no CPU fallback"):
        layer([["a"], ["b"]])


def test_integer_keys_arrive_through_dlpack():
    # any producer that speaks DLPack (tf.experimental.dlpack, CuPy, JAX) can hand integer keys over without a copy
    from recommendflow_b200.backend.layers.preprocess_layers import as_keys

    class Foreign(object):
        def __init__(self, t):
            self._t = t

        def __dlpack__(self, *args, **kwargs):
            return self._t.__dlpack__(*args, **kwargs)

        def __dlpack_device__(self):
            return self._t.__dlpack_device__()
    src = torch.arange(12, dtype=torch.int64).reshape(3, 4)
    got = as_keys(Foreign(src), device="cpu")
    assert got.dtype == torch.int64 and got.shape == (3, 4) and got.data_ptr() == src.data_ptr()
    with pytest.raises(ValueError, match="integer keys"):
        as_keys(Foreign(torch.zeros(2, 2)), device="cpu")
