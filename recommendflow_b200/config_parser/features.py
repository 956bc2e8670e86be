"""Feature specs: `Feature` records and the `Features` collection.

Behavioural mirror of /root/reference/config_parser/features.py (`Feature` :17-89,
`Features` :92-385, `filter_feature`/`except_feature` :388-415): same attribute names and
query helpers, same assertion/exception messages, same quirks (the `...` range expansion
drops the element two before the ellipsis, :224; field-level toggles assert on an empty
name, :357-359).  Checked against the reference's own parser through
tests/golden/config_golden.json.
"""
import os

from .config_proto import (DEFAULT_MAP, PY_CAST, SUPPORT_TYPE, TYPE_MAP, FeatureDeal, FeaturePooling,
                           FeatureTower)
from .config_utils import load_slot_map, load_vocab, read_table
from ..utils.str_parser import str2list

_NO_DIM_DEALS = (FeatureDeal.Numeric, FeatureDeal.Null, FeatureDeal.TokenId, FeatureDeal.Image,
                 FeatureDeal.Embedding, FeatureDeal.BertEncode)


class Feature(object):
    def __init__(self, name, field_name, ftype, tower, deal, vocab_size=-1, embedding_dim=-1,
                 pooling=FeaturePooling("null"), working=True, vocabs=None, seeds=None):
        self.name = name
        self.field_name = field_name
        kind = ftype.lower()
        assert kind in SUPPORT_TYPE, \
            f"Feature type field only support: {SUPPORT_TYPE}, got {ftype}, field: {field_name}"
        self.type = TYPE_MAP[kind]
        self.py_type = kind
        self.tower = tower
        self.deal = deal
        self.vocab_size = vocab_size
        self.embedding_dim = embedding_dim
        self.pooling = pooling
        self.default = DEFAULT_MAP[kind]
        self.working = working
        self.vocabs = [PY_CAST[kind](v) for v in vocabs] if isinstance(vocabs, list) else vocabs
        self.hash_seeds = seeds

    def is_auto_vocabs(self):
        return self.vocabs.upper() == "__AUTO__"

    def is_token_id(self):
        return self.deal == FeatureDeal.TokenId

    def is_lookup(self):
        return self.deal == FeatureDeal.Lookup

    def is_hashing(self):
        return self.deal == FeatureDeal.Hashing

    def is_discrete(self):
        return self.deal == FeatureDeal.Discrete

    def is_image(self):
        return self.deal == FeatureDeal.Image

    def is_embedding(self):
        return self.deal == FeatureDeal.Embedding

    def is_numeric(self):
        return self.deal == FeatureDeal.Numeric

    def is_bert_encode(self):
        return self.deal == FeatureDeal.BertEncode

    # A Feature compares and hashes like its name, so it can index a batch dict directly.
    def __hash__(self):
        return hash(self.name)

    def __eq__(self, other):
        return self.name == (other.name if hasattr(other, "name") else other)

    def __gt__(self, other):
        return self.name > (other.name if hasattr(other, "name") else other)

    def __lt__(self, other):
        return self.name <= other.name if hasattr(other, "name") else self.name < other

    def __repr__(self):
        return (f"Feature(name={self.name!r}, field={self.field_name!r}, type={self.type.name}, "
                f"tower={self.tower.value}, deal={self.deal.value}, vocab_size={self.vocab_size}, "
                f"dim={self.embedding_dim}, pooling={self.pooling.value}, working={self.working})")


def _matches(feature, flag, field, tower, deal, want):
    """Shared body of filter_feature (want=True) / except_feature (want=False); '|' separates terms."""
    checks = []
    if flag:
        checks += [(f in feature.name) for f in flag.split("|")]
    if tower:
        checks += [feature.tower == FeatureTower(t) for t in tower.split("|")]
    if deal:
        checks += [feature.deal == FeatureDeal(d) for d in deal.split("|")]
    if field:
        checks += [feature.field_name == f for f in field.split("|")]
    return all(c == want for c in checks)


def filter_feature(feature, flag=None, field=None, tower=None, deal=None):
    return _matches(feature, flag, field, tower, deal, True)


def except_feature(feature, flag=None, field=None, tower=None, deal=None):
    return _matches(feature, flag, field, tower, deal, False)


class Features(object):
    def __init__(self, conf, vocabs_map=None, seeds=None, slot_map_path=None, normalize_spark_dtypes=False):
        self.conf = conf
        self.slot_map = load_slot_map(slot_map_path, normalize_spark_dtypes) if slot_map_path else {}
        fields = conf["Features"]["feature_fields"]
        self.field_names = fields if isinstance(fields, list) else str2list(fields)
        self.vocabs_map = vocabs_map or {}
        self.seeds = seeds
        self.feature_group = self._init_feature_group(conf["Features"].get("feature_group", {}))
        self.features = self._init_features()
        self._set_attr_by_deal()

    # ---- views -----------------------------------------------------------------------------
    @property
    def train_features(self):
        return [f for f in self.features if f.working]

    @property
    def train_feature_names(self):
        return [f.name for f in self.features if f.working]

    user_features = property(lambda self: self.get_tower_features("user"))
    user_feature_names = property(lambda self: self.get_tower_features("user", True))
    ad_features = property(lambda self: self.get_tower_features("ad"))
    ad_feature_names = property(lambda self: self.get_tower_features("ad", True))
    context_features = property(lambda self: self.get_tower_features("context"))
    context_feature_names = property(lambda self: self.get_tower_features("context", True))
    labels = property(lambda self: self.get_tower_features("label"))
    label_names = property(lambda self: self.get_tower_features("label", True))

    # ---- construction ----------------------------------------------------------------------
    @staticmethod
    def _init_feature_group(groups):
        out = {}
        for key, val in groups.items():
            if isinstance(val, str):
                out[key.lower()] = str2list(val)
            elif isinstance(val, list):
                out[key.lower()] = val
            else:
                raise Exception(f"Feature group except str or list, but got {type(val).__name__}.")
        return out

    def _init_features(self):
        feats, owner = [], {}
        for row in self.conf["Features"]["features"]:
            for feat in self._parse_feature(row):
                if feat.name in owner:
                    raise Exception(
                        f"Feature: [{self.field_names[0]}='{feat.field_name}', name='{feat.name}'] was conflicted with "
                        f"Feature: [{self.field_names[0]}='{owner[feat.name]}', name='{feat.name}']")
                owner[feat.name] = feat.field_name
                feats.append(feat)
        return feats

    def _get_vocab(self, vocab_name, read=True):
        vocab = self.vocabs_map[vocab_name]
        if isinstance(vocab, list):
            return vocab
        if isinstance(vocab, str):
            if not read:
                return vocab
            seen, uniq = set(), []
            for row in read_table(vocab, "\t", ["vocab_id", "vocab_name"]):
                if row[0] not in seen:
                    seen.add(row[0])
                    uniq.append(row[0])
            self.vocabs_map[vocab_name] = uniq
            return uniq
        raise Exception(f"Vocab={vocab_name}, value={vocab}, type={type(vocab)}, expect list or string.")

    def _expand_names(self, field):
        names = self.feature_group[field] if field in self.feature_group else [field]
        if any(isinstance(n, int) for n in names):
            assert self.slot_map, "If you want to set feature slot id to locate feature, you must prepare slot map file."
        while "..." in names:
            at = names.index("...")
            start, end = names[at - 1], names[at + 1]
            assert isinstance(start, int) and isinstance(end, int), f"Except int, got start={start}, end={end}."
            assert start < end, f"Got start={start}, end={end}, start must smaller than end."
            # Reference quirk kept (features.py:224): the slice stops at at-2, so the element two
            # before the ellipsis is dropped ([0, 4, ..., 7] -> [4, 5, 6, 7]).
            names = names[: max(0, at - 2)] + list(range(start, end + 1)) + names[at + 2:]
        for n in names:
            if isinstance(n, int) and n not in self.slot_map:
                raise Exception(f"Feature: [group={field}, slot_id={n}] was not in slot_map_file, please check!")
        return names

    def _parse_feature(self, row):
        d = dict(zip(self.field_names, row))
        assert len(d) == len(self.field_names), f"Conf_str = {row} is invalid, please check."
        field = d[self.field_names[0]].lower()
        names = self._expand_names(field)
        row_type = d["type"].lower()
        name_types = [self.slot_map[n] if isinstance(n, int) else [n, row_type] for n in names]
        tower = FeatureTower(d["tower"].lower())
        deal = FeatureDeal(d["deal"].lower())
        pooling = FeaturePooling(d["pooling"].lower())
        working = d["working"].lower() == "true"
        seeds = self.seeds if deal == FeatureDeal.Hashing else None
        vocab = d["vocab"].lower() if isinstance(d["vocab"], str) else d["vocab"]
        dim = -1 if deal in _NO_DIM_DEALS else int(d["embedding_dim"])

        vocabs, vocab_size = None, -1
        if deal in (FeatureDeal.Lookup, FeatureDeal.Discrete) and working:
            if not isinstance(vocab, str):
                vocabs, vocab_size = vocab, len(vocab)
            elif vocab.startswith("$"):
                vocabs = self._get_vocab(vocab[1:], read=True)
                vocab_size = len(vocabs)
            else:
                try:
                    vocab_size = int(vocab)
                    vocabs = "__AUTO__"
                    assert vocab_size > 0, "Vocab size must be set larger than 0, it means automatically adapt vocabs."
                except ValueError as e:
                    if vocab == "null":
                        raise ValueError("Vocab or vocab size must be given in vocab field when "
                                         "feature deal method set in ['string_lookup', 'integer_lookup', 'discrete']")
                    if vocab in self.vocabs_map:
                        raise Exception(f"Feature field: {field} get vocab symbol: '{vocab}', you may want to set as '${vocab}'?")
                    raise Exception(f"Get unknown vocab symbol: '{vocab}', details: {str(e)}.")
        elif deal == FeatureDeal.BertEncode:
            vocabs = self._get_vocab(vocab[1:], read=False) if vocab.startswith("$") else None
            if vocabs is None:
                raise Exception("Bert encode vocab must given.")
            if not os.path.isfile(vocabs):
                raise FileNotFoundError(f"bert dict vocab path: {vocabs} dose not exist.")
            vocab_size = len(load_vocab(vocabs))
        elif deal == FeatureDeal.Hashing:
            vocab_size = int(vocab)
        return [Feature(n, field, t, tower, deal, vocab_size, dim, pooling, working, vocabs, seeds)
                for n, t in name_types]

    def _set_attr_by_deal(self):
        for deal in FeatureDeal.__members__.values():
            if deal != FeatureDeal.Null:
                setattr(self, f"{deal.value}_features", self.get_deal_features(deal.value))
                setattr(self, f"{deal.value}_feature_names", self.get_deal_features(deal.value, True))

    # ---- queries ---------------------------------------------------------------------------
    def get_tower_features(self, tower, name_only=False):
        want = FeatureTower(tower)
        return [f.name if name_only else f for f in self.train_features if f.tower == want]

    def get_deal_features(self, deal, name_only=False):
        want = FeatureDeal(deal)
        return [f.name if name_only else f for f in self.train_features if f.deal == want]

    def _pool(self, train_only):
        return self.train_features if train_only else self.features

    def _fields_map(self, pred, name_rlike, tower, deal, name_only, train_only):
        out = {}
        for f in self._pool(train_only):
            if pred(f, flag=name_rlike, tower=tower, deal=deal):
                out.setdefault(f.field_name, []).append(f.name if name_only else f)
        return out

    def get_fields_map(self, name_rlike=None, tower=None, deal=None, name_only=False, train_only=True):
        return self._fields_map(filter_feature, name_rlike, tower, deal, name_only, train_only)

    def get_fields_map_except(self, name_rlike=None, tower=None, deal=None, name_only=False, train_only=True):
        return self._fields_map(except_feature, name_rlike, tower, deal, name_only, train_only)

    def get_fields(self, name_rlike=None, tower=None, deal=None, train_only=True):
        return list(self.get_fields_map(name_rlike, tower, deal, True, train_only).keys())

    def get_fields_except(self, name_rlike=None, tower=None, deal=None, train_only=True):
        return list(self.get_fields_map_except(name_rlike, tower, deal, True, train_only).keys())

    def index_of_fields(self, fields_list, name_rlike=None, tower=None, deal=None, train_only=True):
        every = self.get_fields(name_rlike, tower, deal, train_only)
        return [every.index(i) for i in fields_list]

    def get_fields_feature_tuple(self, name_rlike=None, tower=None, deal=None, name_only=False, train_only=True):
        return list(self.get_fields_map(name_rlike, tower, deal, name_only, train_only).values())

    def get_feature(self, name):
        hits = [f for f in self.train_features if f.name == name]
        if not hits:
            raise Exception(f"Feature name = {name} dose not exist.")
        return hits[0]

    def feature_filter(self, name_rlike=None, field=None, tower=None, deal=None, train_only=True):
        return [f for f in list(self._pool(train_only)) if filter_feature(f, name_rlike, field, tower, deal)]

    def feature_except(self, name_rlike=None, field=None, tower=None, deal=None, train_only=True):
        return [f for f in list(self._pool(train_only)) if except_feature(f, name_rlike, field, tower, deal)]

    def get_features(self, name_rlike=None, field=None, tower=None, deal=None, train_only=True):
        return self.feature_filter(name_rlike, field, tower, deal, train_only)

    def index_of_features(self, names, name_rlike=None, field=None, tower=None, deal=None, train_only=True):
        every = [f.name for f in self.feature_filter(name_rlike, field, tower, deal, train_only)]
        return [every.index(n) for n in names]

    def get_features_by_name(self, names=None, prefix="", suffix=""):
        if names:
            return [f for f in self.train_features if f.name in names]
        if prefix:
            return [f for f in self.train_features if f.name.startswith(prefix)]
        if suffix:
            return [f for f in self.train_features if f.name.endswith(suffix)]
        raise ValueError("Names, prefix or suffix must given only one.")

    # ---- toggles ---------------------------------------------------------------------------
    def _set_status(self, name="", field="", status=True):
        assert name or field, "Name or field must given at least one of them"
        # Reference quirk kept (features.py:357-359): the existence check runs on `name` even
        # for field-level toggles, so those assert with an empty feature name.
        assert self.contain(name), f"Feature={name} dose not exists."
        for f in self.features:
            if name and f.name == name:
                f.working = status
            elif field and f.field_name == field:
                f.working = status

    def set_feature_valid(self, name="", field=""):
        self._set_status(name, field, True)

    def set_feature_invalid(self, name="", field=""):
        self._set_status(name, field, False)

    def contain(self, name):
        return any(f.name == name for f in self.train_features)

    def contain_field(self, field):
        return any(f.field_name == field for f in self.train_features)

    def contain_deal(self, deal):
        return any(f.deal == deal for f in self.train_features)

    def get_image_features(self):
        return self.get_deal_features("image")

    def get_embedding_features(self):
        return self.get_deal_features("embedding")
