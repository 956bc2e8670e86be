"""TFRecord / tf.train.Example codec: wire format checked against the protobuf runtime (a dynamically
built tf.train.Example schema) and CRC32C known answers; parse_example densification checked against
the padding rules of the reference's dataloader (backend/core/dataloader.py:23-44)."""
import gzip
import os
import struct

import numpy as np
import pytest

from recommendflow_b200.config_parser import Configuration
from recommendflow_b200.data import tfrecord as tfr


def example_class():
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name="rf_example.proto", package="rf", syntax="proto3")
    for name, typ in (("BytesList", 12), ("FloatList", 2), ("Int64List", 3)):
        m = fd.message_type.add(name=name)
        m.field.add(name="value", number=1, type=typ, label=3)
    feat = fd.message_type.add(name="Feature")
    feat.oneof_decl.add(name="kind")
    for i, t in enumerate(("BytesList", "FloatList", "Int64List")):
        feat.field.add(name=t.lower(), number=i + 1, type=11, label=1, type_name=f".rf.{t}", oneof_index=0)
    feats = fd.message_type.add(name="Features")
    entry = feats.nested_type.add(name="FeatureEntry")
    entry.options.map_entry = True
    entry.field.add(name="key", number=1, type=9, label=1)
    entry.field.add(name="value", number=2, type=11, label=1, type_name=".rf.Feature")
    feats.field.add(name="feature", number=1, type=11, label=3, type_name=".rf.Features.FeatureEntry")
    ex = fd.message_type.add(name="Example")
    ex.field.add(name="features", number=1, type=11, label=1, type_name=".rf.Features")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("rf.Example"))


@pytest.fixture(scope="module")
def conf(golden_dir):
    return Configuration(os.path.join(golden_dir, "configs", "synth_mixed.yaml"))


ROWS = [
    {"clk_items": "i1,i22,i333", "clk_cates": "c9", "uid": "42", "item_id": "it7", "cate_id": "-1", "shop_id": "s1",
     "top_cat": "app", "city_level": "3", "price": "12.5", "avg_price": "3", "dropped": "zz", "label": "1"},
    {"clk_items": "-1", "clk_cates": "c1,c2", "uid": "-7", "item_id": "it8,it9", "cate_id": "k", "shop_id": "s2,s3,s4,s5",
     "top_cat": "game", "city_level": "1,2", "price": "0.25,100", "avg_price": "1e3", "dropped": "-1", "label": "0"},
]


def test_crc32c_known_answers():
    assert tfr.crc32c(b"123456789") == 0xE3069283
    assert tfr.crc32c(b"") == 0
    assert tfr.crc32c(bytes(32)) == 0x8A9136AA          # iSCSI test vector: 32 zero bytes


def test_example_bytes_are_valid_protobuf(conf):
    Example = example_class()
    rec = tfr.build_tfrecord(ROWS[0], conf)
    msg = Example.FromString(rec)
    feats = msg.features.feature
    assert list(feats["clk_items"].byteslist.value) == [b"i1", b"i22", b"i333"]
    assert list(feats["cate_id"].byteslist.value) == [b""]                     # "-1" -> one empty string
    assert list(feats["city_level"].int64list.value) == [3]
    assert list(feats["price"].floatlist.value) == [12.5]
    assert list(feats["label"].floatlist.value) == [1.0]
    assert "dropped" in feats                                                    # non-working features are written too
    # and the other direction: what the protobuf runtime serialises, our decoder reads
    back = tfr.decode_example(msg.SerializeToString())
    assert back["clk_items"] == ("bytes", [b"i1", b"i22", b"i333"]) and back["city_level"] == ("int64", [3])
    neg = tfr.decode_example(tfr.build_tfrecord(ROWS[1], conf))
    assert neg["uid"] == ("bytes", [b"-7"]) and neg["price"][1] == [0.25, 100.0]


def test_tfrecord_file_round_trip_and_padding(conf, tmp_path):
    path = str(tmp_path / "part-0.tfr.gz")
    tfr.dump_tfrecord_data(ROWS, path, conf)
    with gzip.open(path, "rb") as fh:
        raw = fh.read()
    (length,) = struct.unpack("<Q", raw[:8])
    assert struct.unpack("<I", raw[8:12])[0] == tfr.masked_crc32c(raw[:8]) and length == len(tfr.build_tfrecord(ROWS[0], conf))
    recs = list(tfr.read_tfrecord(path, verify_crc=True))
    assert len(recs) == 2
    (batch, labels), = list(tfr.load_tfrecord(path, conf, batch_size=2))
    assert batch["clk_items"].shape == (2, 3)
    assert batch["clk_items"].tolist() == [b"i1", b"i22", b"i333", b"", b"", b""]      # "" pads AND the missing value
    assert batch["shop_id"].shape == (2, 4) and batch["shop_id"].tolist()[:5] == [b"s1", b"", b"", b"", b"s2"]
    assert batch["city_level"].tolist() == [[3, 0], [1, 2]]
    assert np.allclose(batch["price"].numpy(), [[12.5, 0.0], [0.25, 100.0]])
    assert labels["label"].tolist() == [1.0, 0.0] and "dropped" not in batch
    bad = bytearray(raw)
    bad[20] ^= 0xFF
    bad_path = str(tmp_path / "bad.tfr.gz")
    with gzip.open(bad_path, "wb") as fh:
        fh.write(bytes(bad))
    with pytest.raises(IOError):
        list(tfr.read_tfrecord(bad_path, verify_crc=True))
