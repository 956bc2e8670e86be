"""String helpers the config parser needs.

Semantics of /root/reference/utils/str_parser.py:30-44 (`str2list`, `str2dict`) and
:124-144 (`str2loss`, here resolving into recommendflow_b200.backend.lossess).
"""
import importlib

_CASTS = {"str": str, "int": int, "float": float, "set": set, "list": list}


def _cast(kind, value):
    if isinstance(kind, str):
        if kind.lower() not in _CASTS:
            raise ValueError(f"type function: `{kind}` dose not supported")
        return _CASTS[kind.lower()](value)
    return kind(value)


def str2list(input_str, sep=",", trans_type=str):
    """'a, b,,c' -> ['a', 'b', 'c'] (blank items dropped, items stripped)."""
    return [_cast(trans_type, piece.strip()) for piece in input_str.split(sep) if piece.strip()]


def str2dict(input_str, trans_type=str):
    """'a=1;b=2' -> {'a': '1', 'b': '2'}."""
    out = {}
    for item in input_str.strip().split(";"):
        key, value = item.strip().split("=")
        out[key.strip()] = _cast(trans_type, value.strip())
    return out


def _initials(name):
    return "".join(part[0] for part in name.split("_") if part)


def str2loss(loss_name):
    """Resolve a loss by dotted path, bare function name or initials (e.g. 'bnssmccl').

    The reference accepts `backend.losses...` although its package is `backend/lossess`
    (SURVEY.md §5.1); both spellings are accepted here.
    """
    mod_names = ["recommendflow_b200.backend.lossess.match_losses",
                 "recommendflow_b200.backend.lossess.match_zipped_losses"]
    fn_name = loss_name.rsplit(".", 1)[-1]
    if "." in loss_name:
        tail = loss_name.rsplit(".", 2)[-2]
        mod_names = [m for m in mod_names if m.endswith(tail)] or mod_names
    for mod_name in mod_names:
        mod = importlib.import_module(mod_name)
        if hasattr(mod, fn_name):
            return getattr(mod, fn_name)
        for cand in dir(mod):
            if callable(getattr(mod, cand)) and not cand.startswith("_") and _initials(cand) == fn_name:
                return getattr(mod, cand)
    raise ValueError(f"Unknown loss: {loss_name}")
