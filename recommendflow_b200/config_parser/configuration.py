"""YAML configuration with `$var` substitution, feature specs and the experiments table.

Behavioural mirror of /root/reference/config_parser/configuration.py: `Configuration.__init__`
(:25-45), `active_experiment` (:76-102), `get_conf_value` (:104-122), `_set_value`/`_set_str`
(:124-162), `_init_global_conf`/`_rematch_global_conf` (:164-207).  Extensions (opt-in, off by
default so the shipped configs behave exactly as in the reference): `slot_map_path` and
`normalize_spark_dtypes`, which make conf/base_recall_sdpa.yaml usable.
"""
import yaml

from .config_proto import FeatureDeal
from .config_utils import is_punctuation
from .features import Features
from ..utils.str_parser import str2dict, str2list

_SEP = "_##_"


class ExperimentTable(object):
    """The few pandas.DataFrame behaviours the reference uses on `Configuration.experiments`."""

    def __init__(self, rows, columns):
        self.columns = [c for c in columns if c != "exp_id"]
        self._rows = {}
        for row in rows:
            self._rows[row[0]] = dict(zip(columns[1:], row[1:]))   # later duplicates win on .loc

    def __len__(self):
        return len(self._rows)

    @property
    def loc(self):
        return self

    def __getitem__(self, exp_id):
        return _Row(self._rows[exp_id])


class _Row(dict):
    def to_dict(self):
        return dict(self)


class Configuration(object):
    def __init__(self, config_path, slot_map_path=None, normalize_spark_dtypes=False):
        with open(config_path, encoding="utf-8") as fh:
            self.conf = yaml.load(fh.read(), Loader=yaml.FullLoader)
        self._init_global_conf()
        self._rematch_global_conf()
        self.features = Features(self.conf, self.get_conf_value("vocabs"), self.get_conf_value("seeds"),
                                 slot_map_path=slot_map_path, normalize_spark_dtypes=normalize_spark_dtypes)
        self.networks = self.conf["Networks"] if "Networks" in self.conf else {}
        self.exp_conf = self.conf["Experiments"] if "Experiments" in self.conf else None
        if not self.exp_conf or not self.exp_conf["experiments"]:
            self.experiment_field = []
            self.experiments = ExperimentTable([], [])
        else:
            fields = self.exp_conf["experiment_fields"]
            self.experiment_field = str2list(fields) if isinstance(fields, str) else fields
            assert self.experiment_field[0] == "exp_id", "The first field must be exp_id"
            rows = [self._parse_exp(e) for e in self.exp_conf["experiments"]]
            self.experiments = ExperimentTable(rows, self.experiment_field)
        self._refresh_second_parse()

    def _refresh_second_parse(self):
        self.need_parse_second = (self.features.contain_deal(FeatureDeal.Image)
                                  or self.features.contain_deal(FeatureDeal.Embedding))

    @property
    def train_features(self):
        return self.features.train_features

    @property
    def train_feature_names(self):
        return self.features.train_feature_names

    # ---- experiments -----------------------------------------------------------------------
    def _parse_exp(self, cells):
        try:
            exp_id = int(cells[0])
        except Exception as e:
            raise Exception(f"Experiment first col must be integer type exp_id, got {type(cells[0]).__name__}, "
                            f"detail: {str(e)}")
        out = [exp_id]
        for cell in cells[1:]:
            if not isinstance(cell, str):
                out.append(cell)
            elif cell.startswith("{") and cell.endswith("}"):
                out.append(str2dict(cell[1:-1]))
            elif (cell.startswith("[") and cell.endswith("]")) or (cell.startswith("(") and cell.endswith(")")):
                out.append(str2list(cell[1:-1], sep=";"))
            else:
                out.append(self._set_str(cell))
        return out

    def active_experiment(self, exp_id):
        """Apply an experiment's `+name/-name` feature toggles; returns the experiment as a dict."""
        if "features" in self.experiments.columns:
            toggles = self.experiments.loc[exp_id]["features"]
            assert isinstance(toggles, list), "Experiments field features must be a feature name list."
            for item in toggles:
                if item[0] not in "+-":
                    raise ValueError("Feature first latter must be '+/-' represent feature valid/invalid.")
                setter = self.features.set_feature_valid if item[0] == "+" else self.features.set_feature_invalid
                if self.features.contain(item[1:]):
                    setter(name=item[1:])
                else:
                    setter(field=item[1:])
        self._refresh_second_parse()
        return self.experiments.loc[exp_id].to_dict()

    # ---- `$var` machinery ------------------------------------------------------------------
    def get_conf_value(self, key, dtype=None):
        def find(node):
            if key in node:
                return node.get(key)
            for child in node.values():
                if isinstance(child, dict):
                    hit = find(child)
                    if hit is not None:
                        return hit
            return None

        hit = find(self.conf)
        if hit is None:
            raise KeyError(f"Could not find key='{key}' in configuration.")
        return dtype(hit) if dtype else hit

    def _set_value(self, v):
        """A string that is exactly `$name` becomes that value (any type); `$name` inside a longer
        string is spliced in as text."""
        if not isinstance(v, str):
            return v
        bare = not any(is_punctuation(ch, except_char="_$") for ch in v)
        if bare and v.startswith("$"):
            return self.get_conf_value(v[1:])
        if "$" in v:
            return self._set_str(v)
        return v

    def _set_str(self, v):
        if not isinstance(v, str):
            return v
        marked = ""
        for ch in str(v):
            if ch == "$":
                marked += _SEP + "$"
            elif is_punctuation(ch, "_$"):
                marked += _SEP + ch
            else:
                marked += ch
        pieces = []
        for piece in marked.split(_SEP):
            val = self.get_conf_value(piece[1:]) if piece.startswith("$") else piece
            if not isinstance(val, (str, int, float, bool)):
                raise Exception(f"'$' symbol in sub string only support [str, int, float, bool], got {type(val).__name__}. "
                                f"map_value: {val}.")
            pieces.append(str(val))
        return "".join(pieces)

    def _init_global_conf(self):
        feats = self.conf["Features"]
        feats["features"] = [line.split(",") for line in feats["features"].split()]
        exps = self.conf["Experiments"]
        exps["experiments"] = [line.split(",") for line in exps["experiments"].split()] if exps["experiments"] else []

    def _rematch_global_conf(self):
        def walk_list(items):
            out = []
            for it in items:
                if isinstance(it, list):
                    out.append(walk_list(it))
                elif isinstance(it, dict):
                    out.append(walk_dict(it))
                else:
                    out.append(self._set_value(it))
            return out

        def walk_dict(node):
            for k, v in node.items():
                if isinstance(v, dict):
                    walk_dict(v)
                elif isinstance(v, list):
                    new = []
                    for it in v:
                        sub = self._set_value(it)
                        if isinstance(sub, (int, str, float)):
                            new.append(sub)
                        elif isinstance(sub, list):
                            new.append(walk_list(sub))
                        else:
                            raise ValueError(f"'$' symbol in list must be [str, int, float], got {type(sub).__name__}, sub_i: {sub}")
                    node[k] = new
                else:
                    node[k] = self._set_value(v)

        walk_dict(self.conf)

    def print_features(self, scale="train", blank_size=2):
        feats = self.features.features if scale == "all" else self.train_features
        rows = [[f"name={f.name}", f"field={f.field_name}", f"tower={f.tower.value}", f"deal={f.deal.value}",
                 f"type={f.type.name}", f"working={f.working}"] for f in feats]
        widths = [max((len(r[c]) for r in rows), default=0) for c in range(6)]
        for i, r in enumerate(rows):
            body = "".join(cell + " " * (widths[c] - len(cell) + blank_size) for c, cell in enumerate(r[:-1])) + r[-1]
            print(f"Feature {i}:\t[{body}]")
