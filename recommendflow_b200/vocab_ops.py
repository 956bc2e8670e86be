"""Device vocabularies and bucketisation (rf_vocab_* / rf_bucketize_f32): the id-producing front
ends of LookupEmbedding / DiscreteEmbedding (/root/reference/backend/layers/preprocess_layers.py:
134-200).  Term i of the vocabulary maps to id i + 1, anything else to 0 (Keras StringLookup /
IntegerLookup with one OOV index and no mask token).  No CPU fallback."""
import ctypes as C

import numpy as np
import torch

from . import _native as nat
from .strings import StringColumn


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class DeviceVocabulary(object):
    """Open-addressing table in HBM, built by the device from the term list."""

    def __init__(self, terms, device):
        terms = list(terms)
        if len(set(terms)) != len(terms):
            seen, dup = set(), []
            for t in terms:
                if t in seen:
                    dup.append(t)
                seen.add(t)
            raise ValueError(f"The passed vocabulary has at least one repeated term. Please uniquify your dataset. "
                             f"The repeated terms are {dup[:10]}")
        device = torch.device(device)
        if device.type != "cuda":
            raise nat.NativeError("a vocabulary lives on a CUDA device; there is no CPU fallback")
        self.device = device
        self.n_terms = len(terms)
        self.is_int = bool(terms) and all(isinstance(t, (int, np.integer)) for t in terms)
        if terms and not self.is_int and not all(isinstance(t, (str, bytes)) for t in terms):
            raise ValueError("vocabulary terms must be all strings or all integers")
        cap = 2
        while cap < 2 * self.n_terms:
            cap *= 2
        self.slots = torch.empty(cap, dtype=torch.int64, device=device)
        self.desc = nat.VocabDesc(slots=self.slots.data_ptr(), capacity=cap, n_terms=self.n_terms)
        if self.is_int:
            self.term_ints = torch.tensor([int(t) for t in terms], dtype=torch.int64).to(device)
            self.desc.term_ints = self.term_ints.data_ptr()
        elif terms:
            self.terms = StringColumn.from_lists([[t] for t in terms]).to(device)
            self.desc.term_bytes = self.terms.data.data_ptr()
            self.desc.term_offsets = self.terms.offsets.data_ptr()
        with torch.cuda.device(device):
            nat.check(nat.lib().rf_vocab_build(C.byref(self.desc), _stream(device)))

    def lookup(self, keys):
        """StringColumn or int64 tensor on this device -> int64 ids of the same [B, L] shape."""
        with torch.cuda.device(self.device):
            if isinstance(keys, StringColumn):
                out = torch.empty(keys.n_items, dtype=torch.int64, device=self.device)
                nat.check(nat.lib().rf_vocab_lookup_strings(C.byref(self.desc), keys.data.data_ptr(), keys.offsets.data_ptr(),
                                                            keys.n_items, out.data_ptr(), _stream(self.device)))
                return out.view(keys.shape) if keys.shape[1] is not None else out
            if keys.dtype != torch.int64 or not keys.is_cuda:
                raise ValueError("integer keys must be an int64 CUDA tensor")
            vals = keys.contiguous()
            out = torch.empty_like(vals)
            nat.check(nat.lib().rf_vocab_lookup_int64(C.byref(self.desc), vals.data_ptr(), vals.numel(), out.data_ptr(),
                                                      _stream(self.device)))
            return out


def bucketize(values, boundaries):
    """Keras Discretization: id = number of boundaries <= x.  values: fp32 CUDA tensor; boundaries:
    fp32 CUDA tensor, ascending."""
    if not values.is_cuda or not boundaries.is_cuda:
        raise nat.NativeError("bucketize takes CUDA tensors; there is no CPU fallback")
    vals = values.to(torch.float32).contiguous()
    out = torch.empty(vals.shape, dtype=torch.int64, device=vals.device)
    with torch.cuda.device(vals.device):
        nat.check(nat.lib().rf_bucketize_f32(vals.data_ptr(), vals.numel(), boundaries.data_ptr(), boundaries.numel(),
                                             out.data_ptr(), _stream(vals.device)))
    return out
