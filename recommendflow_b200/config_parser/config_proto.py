"""Enums and dtype tags of the feature spec.

Mirrors /root/reference/config_parser/config_proto.py:5-42 (same member names and string
values).  The reference tags feature types with `tf.int64 / tf.float32 / tf.string`;
TensorFlow is not a dependency here, so `DType` carries the same `.name` strings.
"""
from enum import Enum


class DType(object):
    """Stand-in for the tf.DType tags in TYPE_MAP (config_proto.py:41); `.name` matches TF's."""

    def __init__(self, name, torch_name):
        self.name = name
        self.torch_name = torch_name

    def __repr__(self):
        return f"<dtype: '{self.name}'>"

    def __eq__(self, other):
        return getattr(other, "name", other) == self.name

    def __hash__(self):
        return hash(self.name)


int64 = DType("int64", "int64")
float32 = DType("float32", "float32")
string = DType("string", None)


def _spec_enum(name, values):
    """Enum whose member names are the CamelCase of their string values ("token_id" -> TokenId), which is the
    naming the reference's enums follow (config_proto.py:5-33)."""
    members = {"".join(part.capitalize() for part in v.split("_")): v for v in values}
    return Enum(name, members, module=__name__)


FeatureTower = _spec_enum("FeatureTower", ("null", "user", "ad", "context", "label"))
FeatureDeal = _spec_enum("FeatureDeal", ("null", "numeric", "discrete", "hashing", "lookup", "image", "embedding",
                                         "token_id", "bert_encode"))
# NB: no "cls" pooling -- conf/demo_conf.yaml therefore fails to parse, as in the reference.
FeaturePooling = _spec_enum("FeaturePooling", ("null", "avg", "min", "max", "sum", "first", "last"))

TYPE_INT = "int"
TYPE_FLOAT = "float"
TYPE_STR = "str"

SUPPORT_TYPE = [TYPE_INT, TYPE_FLOAT, TYPE_STR]
TYPE_MAP = {TYPE_INT: int64, TYPE_FLOAT: float32, TYPE_STR: string}
DEFAULT_MAP = {TYPE_INT: 0, TYPE_FLOAT: 0.0, TYPE_STR: ""}
PY_CAST = {TYPE_INT: int, TYPE_FLOAT: float, TYPE_STR: str}
