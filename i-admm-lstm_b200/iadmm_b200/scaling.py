"""Drop-in for the reference's `methods/scaling.py` (class `Scaling`), backed by `iadmm_ruiz`.

Same constructor, same `scale_data(Q, p, A0, lb, ub)` return value, same attributes afterwards
(`D, D_inv [B,n,n]`, `E, Einv [B,m,m]`, `c, cinv [B,1,1]`, read by main.py:876-878, :922-940,
:1025-1027).  The dense diagonal matrices are materialised lazily, only if somebody reads them; the
kernels and `LSTM.solve` use the diagonals `d, e, c_vec` directly.
"""
from ctypes import byref, c_size_t

import torch

from . import _lib
from . import ops as _ops          # registers torch.ops.iadmm.*


class Scaling(object):
    def __init__(self, num_var, num_constr, scaling_ites, device):
        self.n = num_var
        self.m = num_constr
        self.device = torch.device(device)
        self.scaling_ites = scaling_ites
        self.MIN_SCALING = 1e-04
        self.MAX_SCALING = 1e04
        self.d = None       # [B, n]  diag(D)
        self.e = None       # [B, m]  diag(E)
        self.c_vec = None   # [B]
        self._dense = {}

    # -- the reference's dense attributes, built on demand -------------------------------------
    def _diag(self, key, vec):
        if vec is None:
            return None
        if key not in self._dense:
            self._dense[key] = torch.diag_embed(vec)
        return self._dense[key]

    @property
    def D(self): return self._diag("D", self.d)
    @property
    def D_inv(self): return self._diag("D_inv", None if self.d is None else torch.reciprocal(self.d))
    @property
    def E(self): return self._diag("E", self.e)
    @property
    def Einv(self): return self._diag("Einv", None if self.e is None else torch.reciprocal(self.e))
    @property
    def c(self): return None if self.c_vec is None else self.c_vec.reshape(-1, 1, 1)
    @property
    def cinv(self): return None if self.c_vec is None else 1.0 / self.c_vec.reshape(-1, 1, 1)

    def scale_data(self, Q, p, A0, lb, ub):
        """methods/scaling.py:50-119.  `lb`/`ub` are the constraint bounds zl/zu (the reference's naming)."""
        L = _lib.lib()
        _lib.require_cuda(Q, p, A0, lb, ub)
        dev = Q.device
        Q, p, A0, zl, zu = (_lib.f32(t, dev) for t in (Q, p, A0, lb, ub))
        B, n, m = Q.shape[0], self.n, self.m
        if Q.shape != (B, n, n) or p.shape[:2] != (B, n) or A0.shape != (B, m, n):
            raise ValueError(f"scale_data: shapes {tuple(Q.shape)}, {tuple(p.shape)}, {tuple(A0.shape)} "
                             f"do not match num_var={n}, num_constr={m}")
        Qs, ps, A0s, zls, zus = (torch.empty_like(t) for t in (Q, p, A0, zl, zu))
        self.d = torch.empty((B, n), dtype=torch.float32, device=dev)
        self.e = torch.empty((B, m), dtype=torch.float32, device=dev)
        self.c_vec = torch.empty((B,), dtype=torch.float32, device=dev)
        self._dense = {}
        nbytes = c_size_t()
        _lib.check(L.iadmm_ruiz_workspace_bytes(B, n, m, byref(nbytes)))
        ws = _lib.workspace(nbytes.value, dev)
        torch.ops.iadmm.ruiz(Q, p, A0, zl, zu, Qs, ps, A0s, zls, zus, self.d, self.e, self.c_vec, ws, int(self.scaling_ites))
        return Qs, ps, A0s, zls, zus
