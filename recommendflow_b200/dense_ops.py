"""Host-side launchers of the dense contractions (rf_sdpa_forward, rf_inbatch_rowstats)."""
import ctypes as C

import torch

from . import _native as nat


# "tf32": tcgen05 tensor cores, fp32 operands read as TF32 (what TensorFlow does for fp32 matmuls on
# Ampere and later GPUs); "fp32": exact CUDA-core accumulation.  Shapes the tensor-core kernels do
# not take use the fp32 kernels.
DEFAULT_PRECISION = "tf32"


def _f32(t, what):
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)
    if not t.is_cuda:
        raise nat.NativeError(f"{what} must live on a CUDA device (got {t.device}); there is no CPU fallback")
    return t.to(torch.float32).contiguous()


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def sdpa_tc_shape_ok(S, dh):
    return 1 <= S <= 64 and dh in (32, 64, 96)


def sdpa(q, k, v, mask=None, precision=None):
    """q, k, v: [..., S, dh]; mask: [..., S, 1] (or [..., S]) of 0/1 -- the reference's query-row mask.
    precision "tf32" runs on the tensor cores when the shape allows (S <= 64, dh in 32/64/96)."""
    precision = precision or DEFAULT_PRECISION
    if precision not in ("tf32", "fp32"):
        raise ValueError("precision must be 'tf32' or 'fp32'")
    q, k, v = _f32(q, "q"), _f32(k, "k"), _f32(v, "v")
    if q.shape != k.shape or q.shape != v.shape:
        raise ValueError(f"q, k, v must share one shape, got {tuple(q.shape)}, {tuple(k.shape)}, {tuple(v.shape)}")
    S, dh = q.shape[-2], q.shape[-1]
    nb = q.numel() // (S * dh) if S * dh else 0
    m = None
    if mask is not None:
        m = _f32(mask, "mask")
        if m.dim() == q.dim() and m.shape[-1] == 1:
            m = m[..., 0]
        m = m.expand(q.shape[:-1]).contiguous()
    out = torch.empty_like(q)
    fn = nat.lib().rf_sdpa_forward_tc if (precision == "tf32" and sdpa_tc_shape_ok(S, dh)) else nat.lib().rf_sdpa_forward
    with torch.cuda.device(q.device):
        nat.check(fn(q.data_ptr(), k.data_ptr(), v.data_ptr(), None if m is None else m.data_ptr(),
                     nb, S, dh, out.data_ptr(), _stream(q.device)))
    return out


def inbatch_rowstats(query, doc, y_true=None, col_weight=None, scale=20.0, margin=0.0, want=("lse", "diag"),
                     precision=None):
    """Row statistics of S = query . doc^T without materialising S.  Returns a dict with the
    requested [B] vectors among lse / diag / hinge / maxoff, plus "loss" when y_true is given."""
    precision = precision or DEFAULT_PRECISION
    if precision not in ("tf32", "fp32"):
        raise ValueError("precision must be 'tf32' or 'fp32'")
    q, d = _f32(query, "query"), _f32(doc, "doc")
    if q.dim() != 2 or q.shape != d.shape:
        raise ValueError(f"query and doc must both be [B, D], got {tuple(q.shape)} and {tuple(d.shape)}")
    B, D = q.shape
    dev = q.device
    y = None if y_true is None else _f32(y_true, "y_true").reshape(-1)
    cw = None if col_weight is None else _f32(col_weight, "col_weight").reshape(-1)
    if y is not None and y.numel() != B:
        raise ValueError("y_true must have one entry per row")
    use_tc = precision == "tf32" and D % 4 == 0
    ws_bytes = nat.lib().rf_inbatch_workspace_bytes_tc(B, D) if use_tc else nat.lib().rf_inbatch_workspace_bytes(B)
    ws = torch.empty(max(1, ws_bytes), dtype=torch.uint8, device=dev)
    res = {k: torch.empty(B, dtype=torch.float32, device=dev) for k in want}
    loss = torch.zeros((), dtype=torch.float32, device=dev) if y is not None else None
    ptr = lambda t: None if t is None else t.data_ptr()
    fn = nat.lib().rf_inbatch_rowstats_tc if use_tc else nat.lib().rf_inbatch_rowstats
    with torch.cuda.device(dev):
        nat.check(fn(q.data_ptr(), d.data_ptr(), ptr(y), ptr(cw), B, D, float(scale), float(margin),
                     ws.data_ptr(), ptr(res.get("lse")), ptr(res.get("diag")), ptr(res.get("hinge")),
                     ptr(res.get("maxoff")), ptr(loss), _stream(dev)))
    if loss is not None:
        res["loss"] = loss
    return res
