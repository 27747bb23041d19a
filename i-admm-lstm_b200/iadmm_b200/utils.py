"""Drop-in for the residual/metric functions of the reference's `utils.py` on the solve path.

`primal_dual_loss` (utils.py:68-71) is one streaming pass over Q and A0 through `iadmm_residuals`.
The adjacent metrics (`obj_fn`, `ineq_dist`, `eq_dist`, utils.py:53-60) are kept as host-side torch
expressions: they are reporting code outside the hot path.
"""
from ctypes import byref, c_size_t

import torch

from . import _lib
from . import ops as _ops          # registers torch.ops.iadmm.*


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
    torch.ops.iadmm.residuals(x, y, z, Q, p, A0, pri, dual, ws)
    pri = pri.reshape(B, 1, 1)
    dual = dual.reshape(B, 1, 1)
    return pri, dual, pri + dual


def _colsum(t):
    return t.sum(dim=1, keepdim=True)


def obj_fn(x, Q, p):
    """0.5 x^T Q x + p^T x per instance, [B,1,1] (utils.py:53-54)."""
    return 0.5 * _colsum(x * torch.bmm(Q, x)) + _colsum(p * x)


def ineq_dist(x, G, c):
    """Violation of G x <= c, [B,mi,1] (utils.py:56-57)."""
    return torch.relu(torch.bmm(G, x) - c)


def eq_dist(x, A, b):
    """Violation of A x = b, [B,me,1] (utils.py:59-60)."""
    return (torch.bmm(A, x) - b).abs()


def lb_dist(x, lb):
    """Violation of x >= lb (utils.py:62-63)."""
    return torch.relu(lb - x)


def ub_dist(x, ub):
    """Violation of x <= ub (utils.py:65-66)."""
    return torch.relu(x - ub)
