"""Drop-in for the reference's `models/lstm.py` (class `LSTM`), backed by libiadmm_b200.

Same constructor, parameter names/shapes (the state_dict contract of models/lstm.py:21-41, so
reference checkpoints load unchanged) and the same `forward(t, num_ineq, num_eq, x, y, z, xv, sigma,
H_t, C_t, **kwargs)` 9-tuple.  `forward` runs ONE iteration through the same CUDA kernels as the fused
`solve`, which runs K iterations and the per-iteration residual evaluation of utils.py:68-71 in one
call with no host synchronisation.
"""
import weakref
from ctypes import byref, c_int, c_size_t
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from . import ops as _ops          # registers torch.ops.iadmm.*

PARAM_ORDER = ("W_i", "U_i", "b_i", "W_f", "U_f", "b_f", "W_o", "U_o", "b_o",
               "W_u", "U_u", "b_u", "W_h", "b_h", "rho", "alpha")


# "auto": a matrix goes to the bitmap-slab form when its densest instance is below this density.  Measured crossover on B200
# (tools/sparse_sweep.py, profiles/r02_sparse_sweep.jsonl, n = m = 1000): the sparse pass is 1.28x faster than the dense one at
# 0.1 % density (diagonal Q of the QP family, identity blocks of SVM, QPLIB-class matrices) but SLOWER from ~0.5 % upwards
# although it reads 3-20x fewer bytes -- ncu: the mask expansion makes the pass instruction-issue bound (91 vs ~25 warp
# instructions per row and 128-column slab), so the 50-60 % dense families stay on the dense streaming pass.
SPARSE_AUTO_DENSITY = 0.003


BLOCK_AUTO_OCCUPANCY = 0.7     # "auto": dense matrix with block skipping when at most this share of its 8x128 blocks is non-empty


def _version(t):
    """`t._version`, or None for inference tensors (`torch.inference_mode()`), which do not track in-place edits: anything keyed
    on a version then simply does not match and the converting / rebuilding path runs."""
    return None if t.is_inference() else t._version


def row_classes(num_ineq, num_eq, m):
    """(inequality rows, equality rows) the way the kernels take them -- the first rows of A0 are inequality rows, the rest
    equality rows, together m = A0.shape[1] -- from the counts main.py hands to `model(t, num_ineq, num_eq, ...)`.  Those
    are the row counts of the `G` and `A` entries of the instance FILE (main.py:248-272), and models/lstm.py:61-62 uses them
    only as the slice `rho_vec[:, num_ineq:num_ineq+num_eq] *= 1e3` of a [B, m, 1] tensor, so they need not add up to m:
    Random_QP files carry G = [A0; -A0] (num_ineq = 2m, generate_data.py:116), SVM files G without the identity rows of A0
    (num_ineq < m, num_eq = 0, :202-207).  Rows outside the slice are inequality-class rows."""
    num_ineq, num_eq, m = int(num_ineq), int(num_eq), int(m)
    if num_ineq < 0 or num_eq < 0:
        raise ValueError(f"negative row count: num_ineq={num_ineq} num_eq={num_eq}")
    a, b = min(num_ineq, m), min(num_ineq + num_eq, m)
    if a == b:
        return m, 0
    if b == m:
        return a, m - a
    raise ValueError(f"rows {a}..{b - 1} of the {m} rows of A0 are equality rows and inequality-class rows follow them: "
                     "the kernels take inequality rows first, then equality rows (generate_data.py:74)")


class SparseBatch:
    """A [B, rows, n] matrix batch in one of the library's two sparse forms (include/iadmm.h):
    kind "slabs"  -- bitmap slabs (iadmm_sparse_pack): masks + packed non-zero values, for unstructured patterns below ~0.3 %;
    kind "blocks" -- the dense matrix plus one occupancy bit per 8-row x 128-column block (iadmm_block_mask): structured
                     sparsity (diagonal, identity blocks, bands) at no decode cost."""

    def __init__(self, kind, buf, cap, shape, nnz=None, nonempty=None):
        self.kind, self.buf, self.cap, self.shape, self.nnz, self.nonempty = kind, buf, cap, tuple(shape), nnz, nonempty

    @property
    def density(self):
        return float(self.nnz.max()) / max(1, self.shape[1] * self.shape[2])

    @property
    def occupancy(self):
        """Share of non-empty 8x128 blocks (worst instance)."""
        total = ((self.shape[1] + 7) // 8) * ((self.shape[2] + 127) // 128)
        return float(self.nonempty.max()) / max(1, total)

    @property
    def bytes_per_instance(self):
        """What one streaming pass reads per instance (worst instance)."""
        if self.kind == "blocks":      # (edge blocks are smaller than 4 KB: never more than the dense matrix)
            return min(4096 * int(self.nonempty.max()), 4 * self.shape[1] * self.shape[2]) + 8 * ((self.shape[1] + 7) // 8)
        return 4 * int(self.nnz.max()) + 20 * self.shape[1] * ((self.shape[2] + 127) // 128)

    @staticmethod
    def pack(M, max_density=None, cap=None):
        """Bitmap-slab form of a dense CUDA batch.  Without `cap` the value capacity is the densest instance's non-zero count
        (one host sync) and None is returned when `max_density` is given and exceeded.  With `cap` (e.g. the capacity of a
        previous batch of the same family, or rows*n) nothing synchronises; check `nnz.max() <= cap` afterwards."""
        _lib.require_cuda(M)
        M = _lib.f32(M)
        B, rows, n = M.shape
        if rows == 0:
            return None
        if cap is None:
            cnt = int(torch.count_nonzero(M.reshape(B, -1), dim=1).max())
            if max_density is not None and cnt > max_density * rows * n:
                return None
            cap = max(4, (cnt + 3) // 4 * 4)
        nbytes = c_size_t()
        _lib.check(_lib.lib().iadmm_sparse_bytes(B, rows, n, cap, byref(nbytes)))
        buf = torch.empty(nbytes.value, dtype=torch.uint8, device=M.device)
        nnz = torch.empty((B,), dtype=torch.int32, device=M.device)
        torch.ops.iadmm.sparse_pack(M, buf, nnz, cap)
        return SparseBatch("slabs", buf, cap, M.shape, nnz=nnz)

    @staticmethod
    def blocks(M, max_occupancy=None):
        """Block-occupancy words of a dense CUDA batch (the matrix itself stays dense).  With `max_occupancy` the share of
        non-empty blocks is read back (one host sync) and None is returned when it is exceeded."""
        _lib.require_cuda(M)
        M = _lib.f32(M)
        B, rows, n = M.shape
        if rows == 0 or n > 8192:
            return None
        nbytes = c_size_t()
        _lib.check(_lib.lib().iadmm_block_mask_bytes(B, rows, n, byref(nbytes)))
        buf = torch.empty(nbytes.value, dtype=torch.uint8, device=M.device)
        nonempty = torch.empty((B,), dtype=torch.int32, device=M.device)
        torch.ops.iadmm.block_mask(M, buf, nonempty)
        sb = SparseBatch("blocks", buf, 0, M.shape, nonempty=nonempty)
        if max_occupancy is not None and sb.occupancy > max_occupancy:
            return None
        return sb

    @staticmethod
    def auto(M):
        """Selection by measured structure / density.  Block skipping first -- it has no decode cost, so it wins whenever at most
        BLOCK_AUTO_OCCUPANCY of the 8x128 blocks are non-empty (measured: diagonal Q at n = 1000 reads 3.5x more bytes as blocks
        than as bitmap slabs and is still ~5x faster); bitmap slabs for unstructured patterns below SPARSE_AUTO_DENSITY; else
        None (plain dense streaming)."""
        sb = SparseBatch.blocks(M, BLOCK_AUTO_OCCUPANCY)
        return sb if sb is not None else SparseBatch.pack(M, SPARSE_AUTO_DENSITY)


@dataclass
class SolveResult:
    x: torch.Tensor            # [B,n,1]
    y: torch.Tensor            # [B,m,1]
    z: torch.Tensor            # [B,m,1]
    xv: torch.Tensor           # [B,n+m,1]
    H: torch.Tensor            # [B,n+m,h]
    C: torch.Tensor            # [B,n+m,h]
    pri: Optional[torch.Tensor] = None           # [K,B] residuals on the data the solve ran on
    dual: Optional[torch.Tensor] = None
    pri_unscaled: Optional[torch.Tensor] = None  # [K,B] on the original data (needs `scaling`)
    dual_unscaled: Optional[torch.Tensor] = None
    metrics: Optional[torch.Tensor] = None       # [K,6,B]: objective, ineq max/mean, eq max/mean, ||K xv - rhs|| (main.py:949-968)

    @property
    def objective(self): return None if self.metrics is None else self.metrics[:, 0]
    @property
    def ineq_violation_max(self): return None if self.metrics is None else self.metrics[:, 1]
    @property
    def ineq_violation_mean(self): return None if self.metrics is None else self.metrics[:, 2]
    @property
    def eq_violation_max(self): return None if self.metrics is None else self.metrics[:, 3]
    @property
    def eq_violation_mean(self): return None if self.metrics is None else self.metrics[:, 4]
    @property
    def ls_residual(self): return None if self.metrics is None else self.metrics[:, 5]     # main.py:952


class LSTM(nn.Module):
    def __init__(self, num_constr, input_dim, hidden_dim, length, device, gate_mode="tc_f16f8"):
        super(LSTM, self).__init__()
        if input_dim != 2:
            raise ValueError("the I-ADMM-LSTM cell takes [xv, grad] (input_dim=2, models/lstm.py:72)")
        self.num_constr = num_constr
        self.input_dim = input_dim
        self.hidden_dim = hidden_dim
        self.length = length
        self.RHO_EQ_OVER_RHO_INEQ = 1e03
        self.device = torch.device(device)
        self.gate_mode = gate_mode
        # forward() returns the dense A_tild like the reference (a fresh [B,N,N] tensor per call, 16 MB per instance at
        # n = m = 1000).  "shared": ONE buffer per (Q, A0, sigma) whose -1/rho_t diagonal is rewritten by every call -- the
        # tensor returned by call t is only valid until call t+1 (main.py re-binds it every iteration and uses the last one).
        # False: A_tild is None.
        self.materialize_kkt = True
        dev = self.device

        def normal(*size):
            return nn.Parameter(torch.normal(mean=0, std=0.01, size=size, device=dev), requires_grad=True)

        def zeros(*size):
            return nn.Parameter(torch.zeros(size, device=dev, dtype=torch.float32), requires_grad=True)

        # same creation order as the reference so a given torch seed draws the same weights
        for g in ("i", "f", "o", "u"):
            setattr(self, f"W_{g}", normal(input_dim, hidden_dim))
            setattr(self, f"U_{g}", normal(hidden_dim, hidden_dim))
            setattr(self, f"b_{g}", zeros(hidden_dim))
        self.W_h = normal(hidden_dim, 1)
        self.b_h = zeros(1)
        self.rho = normal(length, 1)
        self.alpha = normal(length, 1)
        self._packed = None
        self._packed_key = None
        self._ws = None
        self._chain = None               # record of the last `forward` call (see `_step`)
        self.resumed_calls = 0           # how many `forward` calls took the resumed path (diagnostic)
        self._kkt_shared = None          # `materialize_kkt = "shared"`: (key, weakrefs, buffer)

    def name(self):
        return 'lstm'

    _TRANSIENT = ("_packed", "_packed_key", "_ws", "_chain", "_kkt_shared", "_window_ws", "_train_ws")

    def __getstate__(self):
        """Pickling / `copy.deepcopy` / `torch.save(model)` carry the parameters, not the device workspaces, packed weight images
        and per-call records (which hold weak references): those are rebuilt on the next call."""
        state = self.__dict__.copy()
        for k in self._TRANSIENT:
            if k in state:
                state[k] = None
        return state

    # -- packed weights ----------------------------------------------------------------------
    def _mode(self):
        mode = self.gate_mode
        if isinstance(mode, str):
            mode = _lib.GATE_MODES[mode]
        # (hidden_dim % 16 == 8, e.g. configs/QP.yaml's 200, is fine for the fp16+fp8 modes: the solve then always uses the
        # row-interleaved kernels, whose last 16-unit group of e4m3 operands is half zero padding)
        if mode != _lib.GATES_SIMT_FP32 and self.hidden_dim % 8 != 0:
            mode = _lib.GATES_SIMT_FP32    # the tcgen05 tiles need 16-byte rows of fp16
        return mode

    def invalidate_packed(self):
        """Force a re-pack at the next call.  The cache below notices optimizer steps, `load_state_dict`, `copy_` and
        every other in-place op on the parameters (they bump `param._version`); an edit through `param.data`
        (`p.data.add_(...)`, weight surgery in older code) does NOT -- call this after one."""
        self._packed_key = None

    def packed_weights(self):
        """Device buffer in the kernels' layout; re-packed when any parameter changed (see `invalidate_packed`)."""
        prm = [getattr(self, k) for k in PARAM_ORDER]
        key = tuple((p.data_ptr(), p._version) for p in prm)
        if self._packed is None or key != self._packed_key:
            L = _lib.lib()
            dev = prm[0].device
            nbytes = c_size_t()
            _lib.check(L.iadmm_weights_bytes(self.hidden_dim, self.length, byref(nbytes)))
            if self._packed is None or self._packed.numel() != nbytes.value or self._packed.device != dev:
                self._packed = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
            data = [_lib.f32(p.detach()) for p in prm]
            with torch.cuda.device(dev):
                _lib.check(L.iadmm_pack_weights(*[_lib.ptr(t) for t in data], self.hidden_dim, self.length,
                                                _lib.ptr(self._packed), _lib.stream_ptr()))
            self._packed_key = key
        return self._packed

    def _workspace(self, B, n, m, mode, dev):
        nbytes = c_size_t()
        _lib.check(_lib.lib().iadmm_solve_workspace_bytes(B, n, m, self.hidden_dim, mode, byref(nbytes)))
        if self._ws is None or self._ws.numel() < nbytes.value or self._ws.device != dev:
            self._ws = None
            self._ws = _lib.workspace(nbytes.value, dev)
        return self._ws

    # -- K fused iterations -------------------------------------------------------------------
    def solve(self, K, num_ineq, num_eq, Q, p, A0, zl, zu, sigma, state=None, t0=0, scaling=None,
              traces=True, inplace=False, streaming=False, sparse=None):
        """K iterations of `forward` (t = t0..t0+K-1) plus the residuals of utils.py:68-71 after each,
        in one library call.  Small instances (n+m <= 256, hidden_dim 64, tensor-core modes) run on the
        on-chip-resident kernel (one persistent CTA per instance); `streaming=True` forces the HBM-streaming path.  `state=(x,y,z,xv,H,C)` or None for the zero state of main.py:837-843.
        `scaling` is the `Scaling` object that produced (Q,p,A0,zl,zu): with it the residuals of the
        un-scaled iterates on the original data (main.py:922-955) are traced too.
        `sparse`: None = stream Q and A0 dense (the reference densifies every family, main.py:243-296); "auto" = per matrix,
        by measured density / structure (SparseBatch.auto): bitmap slabs for unstructured patterns below 0.3 %, block skipping
        on the dense layout when whole 8x128 blocks are empty (diagonal Q, identity blocks of SVM, banded QPLIB), else dense;
        True / "slabs" = bitmap slabs always, "blocks" = block skipping always; or a pair `(SparseBatch | None, SparseBatch |
        None)` prepared by the caller.  The KKT passes then read only the stored bytes; results are bit-identical to the
        dense path.  One or two host syncs per matrix for "auto"."""
        L = _lib.lib()
        _lib.require_cuda(Q, p, A0, zl, zu, *[prm for prm in self.parameters()])
        self._chain = None               # the workspace is about to be rewritten
        dev = Q.device
        Q, p, A0, zl, zu = (_lib.f32(t, dev) for t in (Q, p, A0, zl, zu))
        B, n = Q.shape[0], Q.shape[1]
        if A0.dim() != 3 or A0.shape[0] != B or A0.shape[2] != n:
            raise ValueError(f"A0 has shape {tuple(A0.shape)}, expected {(B, 'm', n)}")
        m = A0.shape[1]
        num_ineq, num_eq = row_classes(num_ineq, num_eq, m)
        h = self.hidden_dim
        flags = _lib.F_STREAMING if streaming else 0
        if state is None:
            x = torch.zeros((B, n, 1), device=dev); y = torch.zeros((B, m, 1), device=dev)
            z = torch.zeros((B, m, 1), device=dev); xv = torch.zeros((B, n + m, 1), device=dev)
            H = torch.zeros((B, n + m, h), device=dev); C = torch.zeros((B, n + m, h), device=dev)
            flags |= _lib.F_ZERO_STATE
        else:
            x, y, z, xv, H, C = (_lib.f32(t, dev) if inplace else _lib.f32(t, dev).clone() for t in state)
        mode = self._mode()
        packed = self.packed_weights()
        ws = self._workspace(B, n, m, mode, dev)
        pri = dual = pri_u = dual_u = met = None
        if traces and K > 0:
            pri = torch.empty((K, B), device=dev); dual = torch.empty((K, B), device=dev)
            met = torch.empty((K, 6, B), device=dev)
            if scaling is not None:
                pri_u = torch.empty((K, B), device=dev); dual_u = torch.empty((K, B), device=dev)
        sd = se = sc = None
        if scaling is not None:
            sd, se, sc = scaling.d, scaling.e, scaling.c_vec
        for name, v in (("x", x), ("y", y), ("z", z), ("xv", xv), ("H", H), ("C", C)):
            if not v.is_contiguous():
                raise _lib.IadmmError(f"state tensor {name} must be contiguous")
        if sparse is not None and sparse is not False:
            if isinstance(sparse, (tuple, list)):
                q_sp, a_sp = sparse
            elif sparse == "auto":
                q_sp, a_sp = SparseBatch.auto(Q), SparseBatch.auto(A0)
            elif sparse == "blocks":
                q_sp, a_sp = SparseBatch.blocks(Q), SparseBatch.blocks(A0)
            else:
                q_sp, a_sp = SparseBatch.pack(Q), SparseBatch.pack(A0)
            self.last_sparse = (q_sp, a_sp)
            if q_sp is not None or a_sp is not None:
                def form(sb):          # (slab buffer, capacity, block words)
                    if sb is None:
                        return None, 0, None
                    return (sb.buf, sb.cap, None) if sb.kind == "slabs" else (None, 0, sb.buf)
                (q_buf, q_cap, q_blk), (a_buf, a_cap, a_blk) = form(q_sp), form(a_sp)
                torch.ops.iadmm.solve_sparse(packed, Q, q_buf, q_cap, q_blk, p, A0, a_buf, a_cap, a_blk, zl, zu, sd, se, sc,
                                             x, y, z, xv, H, C, pri, dual, pri_u, dual_u, met, ws,
                                             int(num_ineq), int(num_eq), h, self.length, int(t0), int(K), float(sigma), mode, flags)
                return SolveResult(x, y, z, xv, H, C, pri, dual, pri_u, dual_u, met)
        # the one thin custom op of the forward path (iadmm_b200/ops.py -> iadmm_solve of include/iadmm.h)
        torch.ops.iadmm.solve(packed, Q, p, A0, zl, zu, sd, se, sc, x, y, z, xv, H, C, pri, dual, pri_u, dual_u, met, ws,
                              int(num_ineq), int(num_eq), h, self.length, int(t0), int(K), float(sigma), mode, flags)
        return SolveResult(x, y, z, xv, H, C, pri, dual, pri_u, dual_u, met)

    # -- one truncated-BPTT window, forward + backward, in one library call ------------------------
    def train_window(self, TL, num_ineq, num_eq, Q, p, A0, zl, zu, sigma, state, t0=0, loss_scale=None, inplace=False,
                     recompute_gates=None):
        """The body of main.py:336-358 for one window: TL iterations (t = t0..t0+TL-1), the loss
        `sum_t primal_dual_loss(...).mean() * loss_scale` (default 1/TL; main.py uses 1/outer_T) and its gradient
        w.r.t. every parameter, ADDED to `.grad` like `loss.backward()` does.  Returns (loss, state') with
        state' = (x, y, z, xv, H, C) after the window, detached.  Equivalent to running `forward` +
        `primal_dual_loss` under autograd, without a host round trip per iteration.
        `recompute_gates`: True = do not keep the gate activations over the window (the backward re-runs the gate kernel per
        iteration: same gradients bit for bit, 3x less memory, ~15 % more time); None = only when keeping them would not fit
        in the device memory that is free right now."""
        from .autograd import PARAM_ORDER
        L = _lib.lib()
        _lib.require_cuda(Q, p, A0, zl, zu, *[prm for prm in self.parameters()])
        dev = Q.device
        Q, p, A0, zl, zu = (_lib.f32(t, dev) for t in (Q, p, A0, zl, zu))
        B, n = Q.shape[0], Q.shape[1]
        m = A0.shape[1]
        num_ineq, num_eq = row_classes(num_ineq, num_eq, m)
        h = self.hidden_dim
        x, y, z, xv, H, C = (_lib.f32(t.detach(), dev) if inplace else _lib.f32(t.detach(), dev).clone() for t in state)
        count, nbytes = c_size_t(), c_size_t()
        _lib.check(L.iadmm_param_count(h, self.length, byref(count)))
        flags = _lib.TRAIN_RECOMPUTE_GATES if recompute_gates else 0
        _lib.check(L.iadmm_window_workspace_bytes(B, n, m, h, int(TL), flags, byref(nbytes)))
        ws = getattr(self, "_window_ws", None)
        if recompute_gates is None and (ws is None or ws.numel() < nbytes.value):
            free = torch.cuda.mem_get_info(dev)[0] + (ws.numel() if ws is not None and ws.device == dev else 0)
            if nbytes.value > 0.9 * free:
                flags = _lib.TRAIN_RECOMPUTE_GATES
                _lib.check(L.iadmm_window_workspace_bytes(B, n, m, h, int(TL), flags, byref(nbytes)))
        self.last_window_flags = flags
        if ws is None or ws.numel() < nbytes.value or ws.device != dev:
            self._window_ws = None
            ws = self._window_ws = _lib.workspace(nbytes.value, dev)
        flat = torch.empty((count.value,), device=dev)
        loss = torch.empty((1,), device=dev)
        scale = 1.0 / TL if loss_scale is None else float(loss_scale)
        packed = self.packed_weights()
        with torch.cuda.device(dev):
            _lib.check(L.iadmm_train_window(_lib.ptr(packed), _lib.ptr(Q), _lib.ptr(p), _lib.ptr(A0), _lib.ptr(zl), _lib.ptr(zu),
                                            _lib.ptr(x), _lib.ptr(y), _lib.ptr(z), _lib.ptr(xv), _lib.ptr(H), _lib.ptr(C),
                                            _lib.ptr(flat), _lib.ptr(loss), B, n, int(num_ineq), int(num_eq), h, self.length,
                                            int(t0), int(TL), float(sigma), scale, self._mode(), flags, _lib.ptr(ws), ws.numel(),
                                            _lib.stream_ptr()))
        off = 0
        for name in PARAM_ORDER:
            prm = getattr(self, name)
            k = prm.numel()
            g = flat[off:off + k].view(prm.shape)
            prm.grad = g.clone() if prm.grad is None else prm.grad + g
            off += k
        return loss[0], (x, y, z, xv, H, C)

    def _kkt_tuple(self, t, num_ineq, num_eq, Q, p, A0, x, y, z, sigma, dense=True):
        """(A_tild, b_tild, rho_vec) of models/lstm.py:61-69 for the iterate BEFORE iteration t; A_tild is None unless `dense`."""
        dev = Q.device
        B, n = Q.shape[0], Q.shape[1]
        m = A0.shape[1]
        num_ineq, num_eq = row_classes(num_ineq, num_eq, m)
        N = n + m
        shared = dense == "shared"
        b_tild = torch.empty((B, N, 1), device=dev)
        rho_vec = torch.empty((B, m, 1), device=dev)
        Qc, pc, Ac, xc, yc, zc = (_lib.f32(v, dev) for v in (Q, p, A0, x, y, z))
        A_tild, hit = None, False
        if shared:
            # the matrix depends on the iteration only through the -1/rho_t diagonal of its last block
            key = (Qc.data_ptr(), _version(Qc), Ac.data_ptr(), _version(Ac), float(sigma), B, n, m)
            rec = self._kkt_shared
            hit = (rec is not None and rec[0] == key and rec[1]() is not None and rec[2]() is not None
                   and key[1] is not None and key[3] is not None)
            A_tild = rec[3] if hit else torch.empty((B, N, N), device=dev)
            if not hit:
                self._kkt_shared = None          # (drop the old buffer before keeping the new one)
                self._kkt_shared = (key, weakref.ref(Qc), weakref.ref(Ac), A_tild)
        elif dense:
            A_tild = torch.empty((B, N, N), device=dev)
        torch.ops.iadmm.build_kkt(self.packed_weights(), Qc, pc, Ac, xc, yc, zc, None if hit else A_tild, b_tild, rho_vec,
                                  int(num_ineq), int(num_eq), self.hidden_dim, self.length, int(t), float(sigma))
        if hit and m > 0:
            torch.ops.iadmm.kkt_penalty_diagonal(self.packed_weights(), A_tild, n, int(num_ineq), int(num_eq), self.hidden_dim,
                                                 self.length, int(t))
        return A_tild, b_tild, rho_vec

    # -- the reference's per-iteration interface -------------------------------------------------
    def forward(self, t, num_ineq, num_eq, x, y, z, xv, sigma, H_t, C_t, **kwargs):
        """One iteration; returns (x, y, z, xv, H_t, C_t, A_tild, b_tild, rho_vec) like models/lstm.py:96.
        `lb`/`ub` are accepted and ignored, as in the reference (lstm.py:89-90 is commented out)."""
        Q, p, A0, zl, zu = (kwargs[k] for k in ("Q", "p", "A0", "zl", "zu"))
        num_ineq, num_eq = row_classes(num_ineq, num_eq, A0.shape[1])
        if torch.is_grad_enabled() and (any(prm.requires_grad for prm in self.parameters())
                                        or any(v.requires_grad for v in (x, y, z, xv, H_t, C_t))):
            # training (main.py:336-358): one autograd node per iteration, fp32 forward + hand-written backward
            if not (0 <= int(t) < self.length):
                raise IndexError(f"index {t} is out of bounds for dimension 0 with size {self.length}")
            from .autograd import StepFunction, PARAM_ORDER as _ORDER
            _lib.require_cuda(Q, p, A0, zl, zu, x, y, z, xv, H_t, C_t)
            outs = StepFunction.apply(self, int(t), int(num_ineq), int(num_eq), float(sigma), Q, p, A0, zl, zu,
                                      x, y, z, xv, H_t, C_t, *[getattr(self, k) for k in _ORDER])
            # models/lstm.py:96 always returns A_tild, b_tild, rho_vec.  The training loop (main.py:339) discards them, so
            # here they are plain (detached) tensors: b_tild and rho_vec always, the dense A_tild only when
            # `materialize_kkt` asks for it (16 MB per instance at n=m=1000 -- exactly what the tape-free design avoids).
            # NOTE: forward and backward nodes share one workspace (`model._train_ws`): single-stream use only.
            with torch.no_grad():
                kkt = self._kkt_tuple(int(t), num_ineq, num_eq, Q, p, A0, x.detach(), y.detach(), z.detach(), sigma,
                                      dense=self.materialize_kkt)
            return (*outs, *kkt)
        if not (0 <= int(t) < self.length):
            raise IndexError(f"index {t} is out of bounds for dimension 0 with size {self.length}")
        L = _lib.lib()
        _lib.require_cuda(Q, p, A0, zl, zu, x, y, z, xv, H_t, C_t)
        A_tild, b_tild, rho_vec = self._kkt_tuple(int(t), num_ineq, num_eq, Q, p, A0, x, y, z, sigma,
                                                  dense=self.materialize_kkt)
        out = self._step(int(t), int(num_ineq), int(num_eq), Q, p, A0, zl, zu, float(sigma), (x, y, z, xv, H_t, C_t))
        return (*out, A_tild, b_tild, rho_vec)

    def _step(self, t, num_ineq, num_eq, Q, p, A0, zl, zu, sigma, state):
        """One iteration for `forward`, returning new state tensors like models/lstm.py:82-96 (the inputs are left untouched).
        The reference's loops feed the returned H_t, C_t straight back (main.py:874-887, 338-345).  The kernels keep that state
        in row-interleaved fp16 / e4m3 / fp32 operand planes inside the workspace, so when the H_t and C_t passed in ARE the
        tensors the previous call returned (same storage, not modified since: `_version`), the call resumes from the planes
        (IADMM_F_RESUME of include/iadmm.h) instead of converting 2 x [B, n+m, hidden_dim] floats on the way in and cloning
        them for the way out.  Anything else -- a fresh state, an edited H, another batch, a `solve` in between -- takes the
        converting path.  (An edit of H or C through `.data` is not seen, exactly as for `packed_weights`.)"""
        L = _lib.lib()
        dev = Q.device
        B, n = Q.shape[0], Q.shape[1]
        m, h, mode = num_ineq + num_eq, self.hidden_dim, self._mode()
        yes = c_int(0)
        _lib.check(L.iadmm_solve_state_resumable(n, m, h, mode, 0, byref(yes)))
        if not yes.value:
            r = self.solve(1, num_ineq, num_eq, Q, p, A0, zl, zu, sigma, state=state, t0=t, traces=False)
            return r.x, r.y, r.z, r.xv, r.H, r.C
        Q, p, A0, zl, zu = (_lib.f32(v, dev) for v in (Q, p, A0, zl, zu))
        if tuple(A0.shape) != (B, m, n):
            raise ValueError(f"A0 has shape {tuple(A0.shape)}, expected {(B, m, n)}")
        x, y, z, xv = (_lib.f32(v, dev).clone() for v in state[:4])
        H_t, C_t = state[4], state[5]
        packed = self.packed_weights()
        ws = self._workspace(B, n, m, mode, dev)
        dims = (ws.data_ptr(), B, n, m, h, mode, torch.cuda.current_stream(dev).cuda_stream)
        ch, self._chain = self._chain, None

        def same(v, ref, ptr, ver):
            return (ref() is not None and ver is not None and v.data_ptr() == ptr and _version(v) == ver
                    and v.dtype == torch.float32 and tuple(v.shape) == (B, n + m, h) and v.is_contiguous())

        flags = _lib.F_KEEP_PLANES
        if ch is not None and ch["dims"] == dims and same(H_t, *ch["H"]) and same(C_t, *ch["C"]):
            flags |= _lib.F_RESUME | (_lib.F_RESUME_ODD if ch["odd"] else 0)
            odd = not ch["odd"]
            self.resumed_calls += 1
            H = torch.empty((B, n + m, h), device=dev)
            C = torch.empty((B, n + m, h), device=dev)
        else:
            odd = True
            H, C = _lib.f32(H_t, dev).clone(), _lib.f32(C_t, dev).clone()
        torch.ops.iadmm.solve(packed, Q, p, A0, zl, zu, None, None, None, x, y, z, xv, H, C, None, None, None, None, None, ws,
                              num_ineq, num_eq, h, self.length, t, 1, sigma, mode, flags)
        self._chain = {"dims": dims, "odd": odd, "H": (weakref.ref(H), H.data_ptr(), _version(H)),
                       "C": (weakref.ref(C), C.data_ptr(), _version(C))}
        return x, y, z, xv, H, C
