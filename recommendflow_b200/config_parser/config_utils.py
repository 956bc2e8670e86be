"""Slot-map loader and small helpers of the config layer.

Mirrors the parts of /root/reference/config_parser/config_utils.py that the feature parser
uses: `load_slot_map` (:21-33), `is_punctuation` (:85-95), `load_vocab` (:98-107).
`normalize_spark_dtypes=True` is an opt-in extension: the shipped
conf/base_recall_sdpa.feature.map uses Spark type names that the reference rejects.
"""
import unicodedata

from .config_proto import SUPPORT_TYPE

_SPARK_TO_PY = {"stringtype": "str", "integertype": "int", "longtype": "int",
                "floattype": "float", "doubletype": "float"}


def read_table(path, sep, columns):
    """Local-file subset of reference utils/util.py:210 `read_csv`: all cells as str, NA -> "-1"."""
    if path.startswith("hdfs://"):
        raise FileNotFoundError(f"hdfs paths are not reachable from this build: {path}")
    rows = []
    with open(path, encoding="utf-8") as fh:
        for line in fh:
            line = line.rstrip("\n").rstrip("\r")
            if not line:
                continue
            cells = line.split(sep)
            cells = [c if c != "" else "-1" for c in cells]
            if columns is not None:
                cells = (cells + ["-1"] * len(columns))[:len(columns)]
            rows.append(cells)
    return rows


def load_slot_map(slot_map_path, normalize_spark_dtypes=False):
    """`name:dtype:slot` lines -> {slot(int): [name, dtype]}; dtype must be int/float/str."""
    out = {}
    for name, dtype, slot in read_table(slot_map_path, ":", ["name", "dtype", "slot"]):
        dtype = str(dtype).lower()
        if normalize_spark_dtypes:
            inner = dtype
            if inner.startswith("arraytype("):
                inner = inner[len("arraytype("):].split(",")[0]
            dtype = _SPARK_TO_PY.get(inner, dtype)
        assert dtype in SUPPORT_TYPE, f"Unsupported type {dtype}"
        out[int(slot)] = [str(name), dtype]
    return out


def is_punctuation(ch, except_char=""):
    if ch in except_char:
        return False
    code = ord(ch)
    ascii_punct = 33 <= code <= 47 or 58 <= code <= 64 or 91 <= code <= 96 or 123 <= code <= 126
    return ascii_punct or unicodedata.category(ch).startswith("P")


def load_vocab(dict_path, encoding="utf-8"):
    """BERT vocab file -> {token: index}."""
    table = {}
    with open(dict_path, encoding=encoding) as reader:
        for line in reader:
            parts = line.split()
            token = parts[0] if parts else line.strip()
            table[token] = len(table)
    return table
