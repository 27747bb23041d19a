"""Drop-in for the residual/metric functions of the reference's `utils.py` on the solve path.

`primal_dual_loss` (utils.py:68-71) is one streaming pass over Q and A0 through `iadmm_residuals`.
The adjacent metrics (`obj_fn`, `ineq_dist`, `eq_dist`, utils.py:53-60) are kept as host-side torch
expressions: they are reporting code outside the hot path.
"""
from ctypes import byref, c_size_t

import torch

from . import _lib


def primal_dual_loss(x, y, z, Q, p, A0):
    """Returns (||A0 x - z||, ||Q x + p + A0^T y||, sum), each [B,1,1] like the reference."""
    L = _lib.lib()
    _lib.require_cuda(x, y, z, Q, p, A0)
    if torch.is_grad_enabled() and any(v.requires_grad for v in (x, y, z)):
        from .autograd import ResidualFunction        # training loss (main.py:346): differentiable node
        pri, dual = ResidualFunction.apply(x, y, z, Q, p, A0)
        pri, dual = pri.reshape(-1, 1, 1), dual.reshape(-1, 1, 1)
        return pri, dual, pri + dual
    dev = Q.device
    x, y, z, Q, p, A0 = (_lib.f32(t, dev) for t in (x, y, z, Q, p, A0))
    B, n = Q.shape[0], Q.shape[1]
    m = A0.shape[1]
    pri = torch.empty((B,), dtype=torch.float32, device=dev)
    dual = torch.empty((B,), dtype=torch.float32, device=dev)
    nbytes = c_size_t()
    _lib.check(L.iadmm_residuals_workspace_bytes(B, n, m, byref(nbytes)))
    ws = _lib.workspace(nbytes.value, dev)
    with torch.cuda.device(dev):
        _lib.check(L.iadmm_residuals(_lib.ptr(x), _lib.ptr(y), _lib.ptr(z), _lib.ptr(Q), _lib.ptr(p), _lib.ptr(A0),
                                     _lib.ptr(pri), _lib.ptr(dual), B, n, m, _lib.ptr(ws), ws.numel(),
                                     _lib.stream_ptr()))
    pri = pri.reshape(B, 1, 1)
    dual = dual.reshape(B, 1, 1)
    return pri, dual, pri + dual


def obj_fn(x, Q, p):
    return 0.5 * torch.bmm(x.permute(0, 2, 1), torch.bmm(Q, x)) + torch.bmm(p.permute(0, 2, 1), x)


def ineq_dist(x, G, c):
    return torch.clamp(torch.bmm(G, x) - c, 0)


def eq_dist(x, A, b):
    return torch.abs(b - torch.bmm(A, x))


def lb_dist(x, lb):
    return torch.clamp(lb - x, 0)


def ub_dist(x, ub):
    return torch.clamp(x - ub, 0)
