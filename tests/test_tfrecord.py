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
    from recommendflow_b200 import _native as nat
    lib = nat.lib()
    for data in (b"123456789", b"", bytes(32), bytes(range(256)) * 5 + b"xyz"):     # native slicing-by-8 == table loop
        assert lib.rf_crc32c(data, len(data)) == tfr.crc32c(data)
        assert lib.rf_masked_crc32c(data, len(data)) == tfr.masked_crc32c(data)
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
    # the lenient pure-Python decoder (the `uid` feature is int-typed but written as bytes, see the native test below)
    (batch, labels), = list(tfr.load_tfrecord(path, conf, batch_size=2, native=False))
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


# ---- native codec (librf_b200.so: rf_tfrecord_index / rf_example_parse_columns) vs the Python decoder ----------
def _random_rows(rng, n):
    rows = []
    for _ in range(n):
        def seq(prefix, hi, max_n, missing=0.2):
            if rng.uniform() < missing:
                return "-1"
            return ",".join(f"{prefix}{int(rng.integers(0, hi))}" for _ in range(int(rng.integers(1, max_n + 1))))
        rows.append({"clk_items": seq("i", 900, 6), "clk_cates": seq("c", 30, 4), "uid": "7", "item_id": seq("it", 10**6, 1),
                     "cate_id": seq("k", 300, 1), "shop_id": seq("s", 2000, 5),
                     "top_cat": ",".join(["game", "app", "zz"][int(i)] for i in rng.integers(0, 3, size=int(rng.integers(1, 3)))),
                     "city_level": ",".join(str(int(v)) for v in rng.integers(-3, 8, size=int(rng.integers(1, 4)))),
                     "price": ",".join(f"{v:.3f}" for v in rng.uniform(-5, 1500, size=int(rng.integers(1, 3)))),
                     "avg_price": f"{rng.uniform(0, 200):.2f}", "dropped": "zz", "label": str(int(rng.integers(0, 2)))})
    return rows


def _assert_same_batch(a, b):
    assert set(a) == set(b)
    for name in a:
        if hasattr(a[name], "tolist") and hasattr(a[name], "offsets"):
            assert a[name].shape == b[name].shape and a[name].tolist() == b[name].tolist(), name
            assert a[name].data.numel() == b[name].data.numel()
        else:
            assert a[name].dtype == b[name].dtype and a[name].shape == b[name].shape, name
            assert np.array_equal(a[name].numpy(), b[name].numpy(), equal_nan=True), name


def test_native_codec_matches_python_decoder(golden_dir, tmp_path):
    conf2 = Configuration(os.path.join(golden_dir, "configs", "synth_mixed.yaml"))
    for f in conf2.features.features:
        if f.name == "uid":
            f.working = False
    rows = _random_rows(np.random.default_rng(5), 203)
    path = str(tmp_path / "part-0.tfr.gz")
    tfr.dump_tfrecord_data(rows, path, conf2)
    for bs, drop in ((64, False), (50, True), (1, False), (500, False)):
        native = list(tfr.load_tfrecord(path, conf2, batch_size=bs, drop_remainder=drop))
        python = list(tfr.load_tfrecord(path, conf2, batch_size=bs, drop_remainder=drop, native=False))
        assert len(native) == len(python) > 0
        for (nb, nl), (pb, pl) in zip(native, python):
            _assert_same_batch(nb, pb)
            assert set(nl) == set(pl) == {"label"}
    rf = tfr.RecordFile(path, verify_crc=True)
    assert len(rf) == 203
    # the reference's inconsistency, kept: an int-typed hashing feature is WRITTEN as bytes (make_tfrecord.py:104) but
    # DESCRIBED as int64 (dataloader.py:23-44); tf.io.parse_example rejects that, and so does the native codec
    conf3 = Configuration(os.path.join(golden_dir, "configs", "synth_mixed.yaml"))
    with pytest.raises(ValueError, match="feature uid: expected int64_list, found bytes_list"):
        list(tfr.load_tfrecord(path, conf3, batch_size=8))


def test_native_codec_on_protobuf_runtime_output(tmp_path):
    # Examples serialised by the protobuf runtime (its own field order and packing), plus hand-made corner cases:
    # unpacked repeated scalars, a duplicated key (last wins), an unknown field, an entry without a value
    Example = example_class()
    recs = []
    for i in range(40):
        ex = Example()
        f = ex.features.feature
        f["s"].byteslist.value.extend([b"a" * (i % 5), b"", b"xyz"][: i % 4])
        f["n"].int64list.value.extend([i, -i, 2**62, -2**63][: (i % 5)])
        f["x"].floatlist.value.extend([0.5 * i, -1.25][: (i % 3)])
        if i % 7 == 0:
            del f["s"]
        recs.append(ex.SerializeToString())
    var = tfr._varint
    lf = tfr._len_field
    unpacked_ints = b"".join(b"\x08" + var(v) for v in (5, 2**64 - 3))                       # Int64List.value, wire type 0
    unpacked_floats = b"".join(b"\x0d" + struct.pack("<f", v) for v in (1.5, -2.0))          # FloatList.value, wire type 5
    entry = lambda key, feature: lf(1, lf(1, key) + lf(2, feature))
    hand = lf(1, entry(b"n", lf(3, unpacked_ints)) + entry(b"x", lf(2, unpacked_floats)) +
              entry(b"s", lf(1, lf(1, b"old"))) + entry(b"s", lf(1, lf(1, b"new") + lf(1, b"er"))) +
              entry(b"ignored", lf(1, lf(1, b"zz"))) + lf(1, lf(1, b"novalue")) + b"\x10\x07") + b"\x18\x01"
    recs.append(hand)
    path = str(tmp_path / "pb.tfr")
    with open(path, "wb") as fh:
        for rec in recs:
            head = struct.pack("<Q", len(rec))
            fh.write(head + struct.pack("<I", tfr.masked_crc32c(head)) + rec + struct.pack("<I", tfr.masked_crc32c(rec)))
    desc = {"s": tfr.FixedLenSequenceFeature((), tfr.TFString, True, ""), "n": tfr.FixedLenSequenceFeature((), tfr.INT64, True, 0),
            "x": tfr.FixedLenSequenceFeature((), tfr.TFFloat, True, 0.0), "novalue": tfr.FixedLenSequenceFeature((), tfr.TFString, True, ""),
            "absent": tfr.FixedLenFeature((), tfr.TFFloat, 7.5)}
    rf = tfr.RecordFile(path, compression=None, verify_crc=True)
    n_pb = len(rf) - 1
    got = tfr.parse_example_native(rf, 0, n_pb, desc)
    want = tfr.parse_example(list(tfr.read_tfrecord(path, compression=None))[:n_pb], desc)
    _assert_same_batch(got, want)
    assert got["absent"].tolist() == [7.5] * n_pb
    # the hand-made record (the Python decoder does not skip non-length-delimited unknown fields; the native one must)
    hand_out = tfr.parse_example_native(rf, n_pb, 1, desc)
    assert hand_out["s"].tolist() == [b"new", b"er"]                                   # duplicated key: last wins
    assert hand_out["n"].tolist() == [[5, -3]] and hand_out["x"].tolist() == [[1.5, -2.0]]
    assert hand_out["novalue"].shape == (1, 0) and hand_out["absent"].tolist() == [7.5]
    # a flipped payload byte is caught by the CRC check, a truncated file by the framing walk
    raw = bytearray(open(path, "rb").read())
    raw[20] ^= 0x55
    bad = str(tmp_path / "bad.tfr")
    open(bad, "wb").write(bytes(raw))
    with pytest.raises(IOError, match="CRC mismatch in record 0"):
        tfr.RecordFile(bad, compression=None, verify_crc=True)
    open(bad, "wb").write(bytes(raw[:-3]))
    with pytest.raises(IOError, match="truncated"):
        tfr.RecordFile(bad, compression=None)


def test_native_codec_rejects_or_survives_corrupt_examples(golden_dir):
    # bounds: every read in the native decoder is checked against the record's span; mutated records either parse or
    # raise ValueError, they never crash
    conf2 = Configuration(os.path.join(golden_dir, "configs", "synth_mixed.yaml"))
    for f in conf2.features.features:
        if f.name == "uid":
            f.working = False
    rng = np.random.default_rng(11)
    recs = [tfr.build_tfrecord(r, conf2) for r in _random_rows(rng, 6)]
    desc = tfr.build_feature_description(conf2)

    class OneRecord(object):
        pass
    outcomes = {"ok": 0, "rejected": 0}
    for it in range(1500):
        rec = bytearray(recs[it % len(recs)])
        for _ in range(int(rng.integers(1, 4))):
            if len(rec) < 2:
                break
            mode = int(rng.integers(0, 3))
            if mode == 0:
                rec[int(rng.integers(0, len(rec)))] = int(rng.integers(0, 256))
            elif mode == 1:
                del rec[int(rng.integers(1, len(rec))):]
            else:
                a = int(rng.integers(0, len(rec)))
                rec[a:a + int(rng.integers(0, 4))] = bytes(rng.integers(0, 256, size=int(rng.integers(0, 6)), dtype=np.uint8))
        rf = OneRecord()
        rf.data = np.frombuffer(bytes(rec), dtype=np.uint8).copy()
        rf.offsets, rf.lengths = np.array([0], dtype=np.int64), np.array([len(rec)], dtype=np.int64)
        try:
            tfr.parse_example_native(rf, 0, 1, desc)
            outcomes["ok"] += 1
        except ValueError:
            outcomes["rejected"] += 1
    assert outcomes["rejected"] > 100 and outcomes["ok"] > 10


def test_threaded_dataset_yields_every_file_in_order(golden_dir, tmp_path):
    conf2 = Configuration(os.path.join(golden_dir, "configs", "synth_mixed.yaml"))
    for f in conf2.features.features:
        if f.name == "uid":
            f.working = False
    rng = np.random.default_rng(8)
    paths, sizes = [], [37, 5, 64, 1, 20]
    for i, n in enumerate(sizes):
        paths.append(str(tmp_path / f"part-{i}.tfr.gz"))
        tfr.dump_tfrecord_data(_random_rows(rng, n), paths[-1], conf2)
    desc = tfr.build_feature_description(conf2)
    for threads, drop in ((1, False), (3, False), (8, True)):
        got = list(tfr.get_tfrecord_dataset(paths, desc, conf2.features.label_names, 16, thread_num=threads, drop_remainder=drop,
                                            prefetch_buffer_size=2))
        want = [b for p in paths for b in tfr.load_tfrecord(p, conf2, 16, drop_remainder=drop)]
        assert len(got) == len(want) == sum((n // 16) if drop else -(-n // 16) for n in sizes)
        for (gb, gl), (wb, wl) in zip(got, want):
            _assert_same_batch(gb, wb)
            assert np.array_equal(gl["label"].numpy(), wl["label"].numpy())
    # a broken file surfaces as an exception in the consumer, and an abandoned iterator lets the workers stop
    open(paths[2], "wb").write(b"not gzip")
    with pytest.raises(Exception):
        list(tfr.get_tfrecord_dataset(paths, desc, conf2.features.label_names, 16, thread_num=2))
    it = tfr.get_tfrecord_dataset(paths[:2], desc, conf2.features.label_names, 4, thread_num=2, prefetch_buffer_size=1)
    next(it)
    it.close()
