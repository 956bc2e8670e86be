"""Generate tests/golden/config_golden.json by running the REFERENCE's own config parser.

Run in the build container only (needs /root/reference; the GPU box does not have it):
    python tests/golden/make_config_golden.py

`config_parser/` is the one part of the reference that can be imported here: TensorFlow,
tensorflow_io and case_class are absent, so they are replaced by one-line stub modules (the
parser only touches `tf.int64 / tf.float32 / tf.string` as opaque dtype tags and `CaseClass`
as a base class).  For every config we record either the parsed features or the exception
type + message, so tests/test_config_parser.py can check recommendflow_b200.config_parser
against the reference's real behaviour on:
  - the three shipped configs (conf/base_conf.yaml parses; demo_conf.yaml and
    base_recall_sdpa.yaml fail -- SURVEY.md §5.1),
  - the shipped slot map (Spark dtype names -> "Unsupported type" assertion),
  - two synthetic configs under tests/golden/configs/ (lookup/discrete/hashing, experiments,
    slot ids with `...` ranges).
"""
import json
import os
import sys
import types

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def install_stubs():
    class _DT:
        def __init__(self, name):
            self.name = name

        def __repr__(self):
            return f"tf.{self.name}"

    tf = types.ModuleType("tensorflow")
    tf.int64, tf.float32, tf.string = _DT("int64"), _DT("float32"), _DT("string")
    tf.constant = lambda x: x
    tfio = types.ModuleType("tensorflow_io")
    tfio.version = "stub"
    cc_pkg = types.ModuleType("case_class")
    cc = types.ModuleType("case_class.case_class")

    class CaseClass:
        pass

    cc.CaseClass = CaseClass
    cc_pkg.case_class = cc
    for name, mod in [("tensorflow", tf), ("tensorflow_io", tfio), ("case_class", cc_pkg),
                      ("case_class.case_class", cc)]:
        sys.modules[name] = mod
    sys.path.insert(0, REF)


def feature_record(f):
    return {
        "name": f.name, "field_name": f.field_name, "type": f.type.name, "tower": f.tower.value,
        "deal": f.deal.value, "vocab_size": f.vocab_size, "embedding_dim": f.embedding_dim,
        "pooling": f.pooling.value, "default": f.default, "working": f.working, "vocabs": f.vocabs,
        "hash_seeds": f.hash_seeds,
    }


def features_record(feats):
    return {
        "all": [feature_record(f) for f in feats.features],
        "train_feature_names": feats.train_feature_names,
        "user_feature_names": feats.user_feature_names,
        "ad_feature_names": feats.ad_feature_names,
        "label_names": feats.label_names,
        "hashing_feature_names": getattr(feats, "hashing_feature_names"),
        "lookup_feature_names": getattr(feats, "lookup_feature_names"),
        "fields_map_hashing": feats.get_fields_map(deal="hashing", name_only=True),
    }


def capture(fn):
    try:
        return {"ok": fn()}
    except BaseException as e:  # noqa: BLE001 -- we record whatever the reference raises
        return {"error": type(e).__name__, "message": str(e)}


def main():
    install_stubs()
    import contextlib
    import io

    import yaml
    from config_parser.configuration import Configuration
    from config_parser.config_utils import load_slot_map
    from config_parser.features import Features
    from utils.str_parser import str2dict, str2list

    out = {}

    def conf_case(path, exp_ids=()):
        def run():
            conf = Configuration(path)
            rec = {"features": features_record(conf.features),
                   "experiment_field": conf.experiment_field,
                   "experiments": json.loads(conf.experiments.reset_index().to_json(orient="records"))
                   if len(conf.experiments) else [],
                   "need_parse_second": conf.need_parse_second,
                   "conf_values": {k: conf.get_conf_value(k) for k in ("seeds", "batch_size", "task", "data")
                                   if _has(conf, k)},
                   "active": {}}
            for e in exp_ids:
                def act(e=e):
                    active = conf.active_experiment(e)
                    return {"exp": json.loads(json.dumps(active, default=str)),
                            "train_feature_names": conf.train_feature_names}
                rec["active"][str(e)] = capture(act)
            return rec
        return capture(run)

    def _has(conf, k):
        try:
            conf.get_conf_value(k)
            return True
        except KeyError:
            return False

    with contextlib.redirect_stdout(io.StringIO()):
        out["base_conf.yaml"] = conf_case(f"{REF}/conf/base_conf.yaml")
        out["demo_conf.yaml"] = conf_case(f"{REF}/conf/demo_conf.yaml")
        out["base_recall_sdpa.yaml"] = conf_case(f"{REF}/conf/base_recall_sdpa.yaml")
        out["synth_mixed.yaml"] = conf_case(f"{HERE}/configs/synth_mixed.yaml", exp_ids=(1, 2, 0))
        out["shipped_slot_map"] = capture(lambda: load_slot_map(f"{REF}/conf/base_recall_sdpa.feature.map"))
        out["synth_slot_map"] = capture(
            lambda: {str(k): v for k, v in load_slot_map(f"{HERE}/configs/synth_slots.feature.map").items()})

        def slots():
            raw = yaml.load(open(f"{HERE}/configs/synth_slots.yaml").read(), Loader=yaml.FullLoader)
            raw["Features"]["features"] = [line.split(",") for line in raw["Features"]["features"].split()]
            feats = Features(raw, {}, [2022, 2023], slot_map_path=f"{HERE}/configs/synth_slots.feature.map")
            return features_record(feats)
        out["synth_slots.yaml+map"] = capture(slots)

        def slots_nomap():
            raw = yaml.load(open(f"{HERE}/configs/synth_slots.yaml").read(), Loader=yaml.FullLoader)
            raw["Features"]["features"] = [line.split(",") for line in raw["Features"]["features"].split()]
            return features_record(Features(raw, {}, [2022, 2023]))
        out["synth_slots.yaml-nomap"] = capture(slots_nomap)

    out["str2list"] = {s: str2list(s) for s in ["a, b,c", " x ", "", "1,,2"]}
    out["str2dict"] = {s: str2dict(s) for s in ["a=1;b=2", " k = v "]}
    path = os.path.join(HERE, "config_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True, default=str)
    for k, v in out.items():
        print(k, "->", "ok" if "ok" in v else (v.get("error"), str(v.get("message"))[:100]) if isinstance(v, dict) and "error" in v else "dict")


if __name__ == "__main__":
    main()
