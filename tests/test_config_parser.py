"""recommendflow_b200.config_parser vs the REFERENCE's own parser.

The goldens in tests/golden/config_golden.json were produced by importing
/root/reference/config_parser (tests/golden/make_config_golden.py); shipped configs are
copied nowhere -- the three reference yaml files are re-read from /root/reference only when
present (build container), the synthetic ones live under tests/golden/configs/.
"""
import contextlib
import io
import json
import os

import pytest
import yaml

from recommendflow_b200.config_parser import Configuration, Features
from recommendflow_b200.config_parser.config_utils import load_slot_map
from recommendflow_b200.utils.str_parser import str2dict, str2list

REF_CONF = "/root/reference/conf"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF_CONF), reason="reference tree not present on this box")


@pytest.fixture(scope="module")
def golden(golden_dir):
    with open(os.path.join(golden_dir, "config_golden.json")) as f:
        return json.load(f)


def feature_record(f):
    return {"name": f.name, "field_name": f.field_name, "type": f.type.name, "tower": f.tower.value,
            "deal": f.deal.value, "vocab_size": f.vocab_size, "embedding_dim": f.embedding_dim,
            "pooling": f.pooling.value, "default": f.default, "working": f.working, "vocabs": f.vocabs,
            "hash_seeds": f.hash_seeds}


def features_record(feats):
    return {"all": [feature_record(f) for f in feats.features],
            "train_feature_names": feats.train_feature_names,
            "user_feature_names": feats.user_feature_names,
            "ad_feature_names": feats.ad_feature_names,
            "label_names": feats.label_names,
            "hashing_feature_names": feats.hashing_feature_names,
            "lookup_feature_names": feats.lookup_feature_names,
            "fields_map_hashing": feats.get_fields_map(deal="hashing", name_only=True)}


def expect_error(gold, fn):
    with pytest.raises(BaseException) as ei:
        fn()
    assert type(ei.value).__name__ == gold["error"]
    assert str(ei.value) == gold["message"]


@needs_ref
def test_base_conf_matches_reference(golden):
    with contextlib.redirect_stdout(io.StringIO()):
        conf = Configuration(f"{REF_CONF}/base_conf.yaml")
    gold = golden["base_conf.yaml"]["ok"]
    assert features_record(conf.features) == gold["features"]
    assert len(conf.features.features) == 40 and len(conf.train_features) == 7
    app_id = conf.features.get_feature("app_id")
    assert (app_id.vocab_size, app_id.embedding_dim, app_id.pooling.value, app_id.hash_seeds) == (3000, 16, "sum", [2022, 2023])
    assert conf.need_parse_second == gold["need_parse_second"]
    assert conf.experiment_field == gold["experiment_field"]
    for k, v in gold["conf_values"].items():
        assert conf.get_conf_value(k) == v


@needs_ref
@pytest.mark.parametrize("name", ["demo_conf.yaml", "base_recall_sdpa.yaml"])
def test_shipped_configs_fail_like_reference(golden, name):
    expect_error(golden[name], lambda: Configuration(f"{REF_CONF}/{name}"))


@needs_ref
def test_shipped_slot_map_rejected_like_reference(golden):
    expect_error(golden["shipped_slot_map"], lambda: load_slot_map(f"{REF_CONF}/base_recall_sdpa.feature.map"))


@needs_ref
def test_recall_sdpa_usable_with_opt_in_normalisation():
    conf = Configuration(f"{REF_CONF}/base_recall_sdpa.yaml",
                         slot_map_path=f"{REF_CONF}/base_recall_sdpa.feature.map", normalize_spark_dtypes=True)
    hashing = conf.features.hashing_features
    assert len(conf.train_features) == 231 and len(hashing) == 228
    assert len(conf.features.user_features) == 68 and len(conf.features.ad_features) == 160
    assert conf.features.label_names == ["imei", "ad_id", "label"]
    assert all((f.vocab_size, f.embedding_dim, f.pooling.value) == (100000, 8, "sum") for f in hashing)
    assert "channel" not in conf.train_feature_names      # dropped by the `...` expansion quirk


def test_synth_mixed_matches_reference(golden, golden_dir):
    gold = golden["synth_mixed.yaml"]["ok"]
    conf = Configuration(os.path.join(golden_dir, "configs", "synth_mixed.yaml"))
    assert features_record(conf.features) == gold["features"]
    assert conf.experiment_field == gold["experiment_field"]
    for k, v in gold["conf_values"].items():
        assert conf.get_conf_value(k) == v
    for row in gold["experiments"]:
        got = conf.experiments.loc[row["exp_id"]].to_dict()
        assert got == {k: v for k, v in row.items() if k != "exp_id"}
    for exp_id in ("1", "2", "0"):                         # same order as the generator
        g = gold["active"][exp_id]
        if "ok" in g:
            active = conf.active_experiment(int(exp_id))
            assert json.loads(json.dumps(active, default=str)) == g["ok"]["exp"]
            assert conf.train_feature_names == g["ok"]["train_feature_names"]
        else:
            expect_error(g, lambda: conf.active_experiment(int(exp_id)))


def _raw_slots(golden_dir):
    raw = yaml.load(open(os.path.join(golden_dir, "configs", "synth_slots.yaml")).read(), Loader=yaml.FullLoader)
    raw["Features"]["features"] = [line.split(",") for line in raw["Features"]["features"].split()]
    return raw


def test_slot_ids_and_ellipsis_match_reference(golden, golden_dir):
    smap = os.path.join(golden_dir, "configs", "synth_slots.feature.map")
    assert {str(k): v for k, v in load_slot_map(smap).items()} == golden["synth_slot_map"]["ok"]
    feats = Features(_raw_slots(golden_dir), {}, [2022, 2023], slot_map_path=smap)
    assert features_record(feats) == golden["synth_slots.yaml+map"]["ok"]
    assert "t_17" not in feats.train_feature_names and "channel" not in feats.train_feature_names
    expect_error(golden["synth_slots.yaml-nomap"], lambda: Features(_raw_slots(golden_dir), {}, [2022, 2023]))


def test_str_helpers_match_reference(golden):
    for s, want in golden["str2list"].items():
        assert str2list(s) == want
    for s, want in golden["str2dict"].items():
        assert str2dict(s) == want


def test_feature_behaves_like_its_name(golden_dir):
    conf = Configuration(os.path.join(golden_dir, "configs", "synth_mixed.yaml"))
    uid = conf.features.get_feature("uid")
    batch = {"uid": [1, 2, 3]}
    assert batch[uid] == [1, 2, 3] and uid == "uid" and uid.is_hashing()
    assert conf.features.get_fields(deal="hashing") == ["user_seq", "uid", "ad_ids"]
    assert conf.features.index_of_features(["item_id"], tower="ad") == [0]
