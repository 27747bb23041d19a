"""CPU oracle for the I-ADMM-LSTM unrolled solve path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain torch-on-CPU tensor arithmetic, the algorithm of the one hot path this
repository accelerates.  It is the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  Nothing under
``i-admm-lstm_b200/`` imports it and the product path has no CPU fallback.

Parity pinning: the reference ships no tests, golden vectors, checkpoints or datasets (SURVEY.md
section 4).  The oracle is therefore pinned against outputs of the reference's own modules executed
in the build container: ``tests/golden/make_golden.py`` imports ``/root/reference`` (models/lstm.py,
methods/scaling.py, utils.py), runs them on seeded inputs and commits the in/out tensors as
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function below against them.

Reference lines each function follows (paths relative to the reference checkout):

=====================  ==========================================================
``qp_instances``        generate_data.py:67-76 (QP family), main.py:718 (Q doubled)
``lstm_parameters``     models/lstm.py:21-41 (16 tensors, N(0, 0.01) / zeros)
``penalty_schedule``    models/lstm.py:60-63
``kkt_system``          models/lstm.py:67-69
``lstm_step``           models/lstm.py:47-96
``ruiz_equilibrate``    methods/scaling.py:17-46 and :50-119
``primal_dual_residuals``  utils.py:68-71
``objective`` ...       utils.py:53-66
``unscale_iterates``    main.py:922,923,940 and :1025-1027
``solve``               main.py:837-843 and :874-887 (+ :346 / :955 residual traces)
``exact_admm_step``     models/lu.py:13-47 (Stage II, used as the objective-gap oracle)
=====================  ==========================================================

All functions are dtype-generic: run them in float32 for parity with the reference, in float64 as
the tie-breaker.  Vectors keep the reference's ``[B, dim, 1]`` column layout.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch

RHO_EQ_OVER_RHO_INEQ = 1e3      # models/lstm.py:18
MIN_SCALING = 1e-4              # methods/scaling.py:12
MAX_SCALING = 1e4               # methods/scaling.py:13
GATES = ("i", "f", "o", "u")    # input, forget, output, candidate


# --------------------------------------------------------------------------------------------
# problem data and parameters
# --------------------------------------------------------------------------------------------
def qp_instances(batch: int, n: int, num_ineq: int, num_eq: int, seed: int = 17,
                 dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Synthetic dense QPs of the reference's ``QP`` family.

    generate_data.py:67-76: ``Q0 = 0.5*diag(U[0,1))``, ``p ~ U[0,1)``, ``A ~ N(0,1)``, ``b ~ U[-1,1)``,
    ``G ~ N(0,1)``, ``c = sum_j |G A^+|_ij`` (feasible by construction), rows of ``A0`` are the
    inequality rows first, then the equality rows; ``zl = [-inf; b]``, ``zu = [c; b]``.
    main.py:718 doubles Q when a dataset is loaded, so ``Q = diag(U[0,1))`` here.  The OSQP
    "solved" filter of the reference is skipped (osqp is not installed in this image).
    The reference seeds nothing; we draw from a seeded generator in float32 like the reference.
    """
    g = torch.Generator().manual_seed(seed)
    diag = torch.rand((batch, n), generator=g)
    Q = 2.0 * (0.5 * torch.diag_embed(diag))
    p = torch.rand((batch, n), generator=g).unsqueeze(-1)
    A = torch.randn((batch, num_eq, n), generator=g)
    b = (2.0 * torch.rand((batch, num_eq), generator=g) - 1.0).unsqueeze(-1)
    G = torch.randn((batch, num_ineq, n), generator=g)
    if num_eq > 0 and num_ineq > 0:
        c = torch.sum(torch.abs(torch.bmm(G, torch.pinverse(A))), dim=2).unsqueeze(-1)
    else:
        c = torch.rand((batch, num_ineq), generator=g).unsqueeze(-1) + 1.0
    A0 = torch.cat((G, A), dim=1)
    zl = torch.cat((torch.full_like(c, -math.inf), b), dim=1)
    zu = torch.cat((c, b), dim=1)
    out = dict(Q=Q, p=p, A0=A0, zl=zl, zu=zu, G=G, c=c, A=A, b=b)
    return {k: v.to(dtype).contiguous() for k, v in out.items()}


PARAM_ORDER = ("W_i", "U_i", "b_i", "W_f", "U_f", "b_f", "W_o", "U_o", "b_o",
               "W_u", "U_u", "b_u", "W_h", "b_h", "rho", "alpha")


def lstm_parameters(hidden: int, length: int, seed: int = 17, input_dim: int = 2,
                    scale: float = 1.0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """The 16 learnable tensors of the reference optimiser cell (models/lstm.py:21-41).

    Keys and shapes are the ``state_dict`` contract: ``W_g:[input_dim,h]``, ``U_g:[h,h]``,
    ``b_g:[h]`` for g in i,f,o,u; ``W_h:[h,1]``, ``b_h:[1]``, ``rho, alpha:[length,1]``.
    ``scale`` multiplies the N(0, 0.01) draws (used by tests to emulate trained-size weights).
    """
    g = torch.Generator().manual_seed(seed)
    prm: Dict[str, torch.Tensor] = {}
    for gate in GATES:
        prm[f"W_{gate}"] = 0.01 * scale * torch.randn((input_dim, hidden), generator=g)
        prm[f"U_{gate}"] = 0.01 * scale * torch.randn((hidden, hidden), generator=g)
        prm[f"b_{gate}"] = torch.zeros(hidden)
    prm["W_h"] = 0.01 * scale * torch.randn((hidden, 1), generator=g)
    prm["b_h"] = torch.zeros(1)
    prm["rho"] = 0.01 * torch.randn((length, 1), generator=g)
    prm["alpha"] = 0.01 * torch.randn((length, 1), generator=g)
    return {k: v.to(dtype) for k, v in prm.items()}


# --------------------------------------------------------------------------------------------
# one unrolled iteration
# --------------------------------------------------------------------------------------------
def penalty_schedule(prm, t: int, num_ineq: int, num_eq: int, batch: int, rows: Optional[int] = None):
    """models/lstm.py:60-63: rho_t = sigmoid(rho[t]); equality rows carry 1e3*rho_t;
    alpha_t = 2*sigmoid(alpha[t]).  Returns (rho_vec [B,m,1], alpha [1]).
    The reference sizes rho_vec by ``y.shape`` (lstm.py:61) and only SLICES it with the counts (:62); ``rows`` is that
    row count (default: the counts add up to it, as for the QP family).  main.py takes the counts from the G / A entries
    of the instance file (:248-272), which for Random_QP and SVM files do not add up to the rows of A0."""
    rho = torch.sigmoid(prm["rho"][t, :])
    m = num_ineq + num_eq if rows is None else rows
    rho_vec = torch.ones((batch, m, 1), dtype=rho.dtype) * rho
    rho_vec[:, num_ineq:num_ineq + num_eq, :] = rho_vec[:, num_ineq:num_ineq + num_eq, :] * RHO_EQ_OVER_RHO_INEQ
    alpha = 2 * torch.sigmoid(prm["alpha"][t, :])
    return rho_vec, alpha


def kkt_system(Q, p, A0, x, y, z, rho_vec, sigma: float):
    """models/lstm.py:67-69: the quasi-definite KKT matrix and right-hand side of the x/v subproblem,
    ``K = [[Q + sigma I, A0^T], [A0, -diag(1/rho)]]``, ``rhs = [sigma x - p ; z - y/rho]``."""
    B, n, _ = Q.shape
    m = A0.shape[1]
    eye_n = torch.diag_embed(torch.ones((B, n), dtype=Q.dtype))
    eye_m = torch.diag_embed(torch.ones((B, m), dtype=Q.dtype))
    top = torch.cat((Q + sigma * eye_n, A0.transpose(1, 2)), dim=2)
    bot = torch.cat((A0, -(1 / rho_vec) * eye_m), dim=2)
    K = torch.cat((top, bot), dim=1)
    rhs = torch.cat((sigma * x - p, z - (1 / rho_vec) * y), dim=1)
    return K, rhs


def ls_gradient(Q, p, A0, x, y, z, xv, rho_vec, sigma: float, form: str = "dense"):
    """models/lstm.py:72: gradient of 0.5*||K xv - rhs||^2 w.r.t. xv, i.e. ``K^T (K xv - rhs)``.

    ``form='dense'`` builds K like the reference does (this is what the reference costs on a CPU);
    ``form='block'`` evaluates the same expression block-wise without ever forming K (this is the
    formulation the CUDA kernels use)."""
    n = Q.shape[1]
    if form == "dense":
        K, rhs = kkt_system(Q, p, A0, x, y, z, rho_vec, sigma)
        return torch.bmm(K.transpose(1, 2), torch.bmm(K, xv) - rhs)
    xt, v = xv[:, :n, :], xv[:, n:, :]
    inv_rho = 1 / rho_vec
    w1 = torch.bmm(Q, xt) + sigma * xt + torch.bmm(A0.transpose(1, 2), v) - (sigma * x - p)
    w2 = torch.bmm(A0, xt) - inv_rho * v - (z - inv_rho * y)
    g1 = torch.bmm(Q.transpose(1, 2), w1) + sigma * w1 + torch.bmm(A0.transpose(1, 2), w2)
    g2 = torch.bmm(A0, w1) - inv_rho * w2
    return torch.cat((g1, g2), dim=1)


def lstm_cell(prm, feats, H, C):
    """models/lstm.py:74-80: coordinate-wise LSTM cell with shared weights and the scalar output head.
    ``feats:[B,N,2]`` (current xv and its least-squares gradient), ``H,C:[B,N,h]``."""
    pre = {g: feats @ prm[f"W_{g}"] + H @ prm[f"U_{g}"] + prm[f"b_{g}"] for g in GATES}
    gate_i = torch.sigmoid(pre["i"])
    gate_f = torch.sigmoid(pre["f"])
    gate_o = torch.sigmoid(pre["o"])
    cand = torch.tanh(pre["u"])
    C = gate_i * cand + gate_f * C
    H = gate_o * torch.tanh(C)
    step = H @ prm["W_h"] + prm["b_h"]
    return H, C, step


def admm_update(x, y, z, xv, rho_vec, alpha, zl, zu):
    """models/lstm.py:84-94: OSQP-style x relaxation, z projection onto [zl, zu] and dual update.
    No relaxation on z (lstm.py:91-92); the x box clip of lstm.py:89-90 is disabled upstream."""
    n = x.shape[1]
    xt, v = xv[:, :n, :], xv[:, n:, :]
    z_mid = z + (1 / rho_vec) * (v - y)
    x = alpha * xt + (1 - alpha) * x
    z = torch.max(torch.min(z_mid + (1 / rho_vec) * y, zu), zl)
    y = y + rho_vec * (z_mid - z)
    return x, y, z


def lstm_step(prm, t: int, num_ineq: int, num_eq: int, x, y, z, xv, sigma: float, H, C,
              Q, p, A0, zl, zu, form: str = "dense"):
    """One unrolled I-ADMM-LSTM iteration (models/lstm.py:47-96).
    Returns ``(x, y, z, xv, H, C, rho_vec)``; K and rhs of the reference's 9-tuple are available
    from :func:`kkt_system`."""
    rho_vec, alpha = penalty_schedule(prm, t, num_ineq, num_eq, x.shape[0], rows=y.shape[1])
    grad = ls_gradient(Q, p, A0, x, y, z, xv, rho_vec, sigma, form)
    feats = torch.cat((xv, grad), dim=-1)
    H, C, step = lstm_cell(prm, feats, H, C)
    xv = xv - step
    x, y, z = admm_update(x, y, z, xv, rho_vec, alpha, zl, zu)
    return x, y, z, xv, H, C, rho_vec


# --------------------------------------------------------------------------------------------
# residuals and metrics
# --------------------------------------------------------------------------------------------
def primal_dual_residuals(x, y, z, Q, p, A0):
    """utils.py:68-71: per-instance ||A0 x - z||_2 and ||Q x + p + A0^T y||_2, each [B,1,1]."""
    pri = torch.linalg.vector_norm(torch.bmm(A0, x) - z, dim=(1, 2), keepdim=True)
    dual = torch.linalg.vector_norm(torch.bmm(Q, x) + p + torch.bmm(A0.transpose(1, 2), y),
                                    dim=(1, 2), keepdim=True)
    return pri, dual, pri + dual


def objective(x, Q, p):
    """utils.py:53-54."""
    return 0.5 * torch.bmm(x.transpose(1, 2), torch.bmm(Q, x)) + torch.bmm(p.transpose(1, 2), x)


def ineq_violation(x, G, c):
    """utils.py:56-57."""
    return torch.clamp(torch.bmm(G, x) - c, 0)


def eq_violation(x, A, b):
    """utils.py:59-60."""
    return torch.abs(b - torch.bmm(A, x))


# --------------------------------------------------------------------------------------------
# Ruiz equilibration
# --------------------------------------------------------------------------------------------
def _limit(v):
    """methods/scaling.py:31-46 (tensor branch): clamp to [1e-4, 1e4]; entries that end up exactly
    at the lower clamp become 1 (a zero row/column is left unscaled)."""
    lo = torch.tensor(MIN_SCALING, dtype=v.dtype)
    hi = torch.tensor(MAX_SCALING, dtype=v.dtype)
    w = torch.minimum(torch.maximum(v, lo), hi)
    return torch.where(w == lo, torch.ones_like(w), w)


@dataclass
class RuizScaling:
    """Diagonals of the equilibration (the reference keeps them as dense diag matrices
    ``D, D_inv [B,n,n]``, ``E, Einv [B,m,m]`` and ``c, cinv [B,1,1]``; methods/scaling.py:107-117)."""
    d: torch.Tensor          # [B, n]
    e: torch.Tensor          # [B, m]
    c: torch.Tensor          # [B, 1, 1]

    @property
    def D(self): return torch.diag_embed(self.d)
    @property
    def D_inv(self): return torch.diag_embed(torch.reciprocal(self.d))
    @property
    def E(self): return torch.diag_embed(self.e)
    @property
    def Einv(self): return torch.diag_embed(torch.reciprocal(self.e))
    @property
    def cinv(self): return 1.0 / self.c


def ruiz_equilibrate(Q, p, A0, zl, zu, iterations: int = 10):
    """Modified Ruiz equilibration with cost normalisation (methods/scaling.py:50-119).

    Works on the diagonals only: because every ``bmm`` of the reference has a diagonal factor, each
    entry of its result is a single rounded product, so ``Q <- d_i*(Q_ij*d_j)`` evaluated
    element-wise reproduces the reference's dense-``bmm`` arithmetic exactly, at O(n^2) instead of
    O(n^3).  Returns ``(Q, p, A0, zl, zu, RuizScaling)``.
    """
    B, n, _ = Q.shape
    m = A0.shape[1]
    d = torch.ones((B, n), dtype=Q.dtype)
    e = torch.ones((B, m), dtype=Q.dtype)
    c = torch.ones((B, 1, 1), dtype=Q.dtype)
    for _ in range(iterations):
        # scaling.py:17-29  inf-norms of the KKT columns
        col = torch.maximum(Q.abs().amax(dim=1), A0.abs().amax(dim=1)) if m > 0 else Q.abs().amax(dim=1)
        row = A0.abs().amax(dim=2) if m > 0 else Q.new_zeros((B, 0))
        s = torch.reciprocal(torch.sqrt(_limit(torch.cat((col, row), dim=-1))))   # :66-69
        sd, se = s[:, :n], s[:, n:]
        Q = sd.unsqueeze(2) * (Q * sd.unsqueeze(1))                               # :80  D (Q D)
        A0 = se.unsqueeze(2) * (A0 * sd.unsqueeze(1))                             # :81  E (A0 D)
        p = sd.unsqueeze(2) * p                                                   # :82
        zl = se.unsqueeze(2) * zl                                                 # :83
        zu = se.unsqueeze(2) * zu                                                 # :84
        d = sd * d                                                                # :87
        e = se * e                                                                # :88
        # scaling.py:91-105  cost normalisation
        mean_col = Q.abs().amax(dim=1).mean(-1, keepdim=True)                     # [B,1]
        p_inf = _limit(p.abs().amax(dim=1))                                       # [B,1]
        c_t = 1.0 / _limit(torch.maximum(p_inf, mean_col))                        # [B,1]
        Q = c_t.unsqueeze(-1) * Q
        p = c_t.unsqueeze(-1) * p
        c = c_t.unsqueeze(-1) * c
    return Q, p, A0, zl, zu, RuizScaling(d=d, e=e, c=c)


def unscale_iterates(x, y, z, sc: RuizScaling):
    """main.py:922,923,940: x = D x_s, z = E^-1 z_s, y = c^-1 E y_s (diagonal, so element-wise)."""
    return (sc.d.unsqueeze(-1) * x,
            (sc.cinv.squeeze(-1) * sc.e).unsqueeze(-1) * y,
            torch.reciprocal(sc.e).unsqueeze(-1) * z)


# --------------------------------------------------------------------------------------------
# the unrolled solve
# --------------------------------------------------------------------------------------------
@dataclass
class SolveResult:
    x: torch.Tensor
    y: torch.Tensor
    z: torch.Tensor
    xv: torch.Tensor
    H: torch.Tensor
    C: torch.Tensor
    pri: torch.Tensor                        # [K, B] residuals on the data the solve ran on
    dual: torch.Tensor
    pri_unscaled: Optional[torch.Tensor] = None   # [K, B] on the original data (test mode, main.py:955)
    dual_unscaled: Optional[torch.Tensor] = None
    obj_unscaled: Optional[torch.Tensor] = None   # [K, B] (main.py:950)
    extra: dict = field(default_factory=dict)


def solve(prm, K: int, num_ineq: int, num_eq: int, Q, p, A0, zl, zu, sigma: float, hidden: int,
          t0: int = 0, state=None, scaling: Optional[RuizScaling] = None, original=None,
          form: str = "dense") -> SolveResult:
    """K unrolled iterations from the zero state (main.py:837-843, :874-887).

    ``pri/dual`` are the residuals of utils.py:68-71 on the data passed in (training loss,
    main.py:346).  When ``scaling`` and ``original=(Q,p,A0)`` are given, the test-mode metrics on the
    unscaled iterates and original data are traced as well (main.py:922-955)."""
    B, n, _ = Q.shape
    m = A0.shape[1]
    dt = Q.dtype
    if state is None:
        x = torch.zeros((B, n, 1), dtype=dt)
        y = torch.zeros((B, m, 1), dtype=dt)
        z = torch.zeros((B, m, 1), dtype=dt)
        xv = torch.zeros((B, n + m, 1), dtype=dt)
        H = torch.zeros((B, n + m, hidden), dtype=dt)
        C = torch.zeros((B, n + m, hidden), dtype=dt)
    else:
        x, y, z, xv, H, C = state
    pri_t, dual_t, pri_u, dual_u, obj_u = [], [], [], [], []
    for t in range(t0, t0 + K):
        x, y, z, xv, H, C, _ = lstm_step(prm, t, num_ineq, num_eq, x, y, z, xv, sigma, H, C,
                                         Q, p, A0, zl, zu, form=form)
        pr, du, _ = primal_dual_residuals(x, y, z, Q, p, A0)
        pri_t.append(pr.reshape(B)); dual_t.append(du.reshape(B))
        if scaling is not None and original is not None:
            Qo, po, Ao = original
            xu, yu, zu_ = unscale_iterates(x, y, z, scaling)
            pr, du, _ = primal_dual_residuals(xu, yu, zu_, Qo, po, Ao)
            pri_u.append(pr.reshape(B)); dual_u.append(du.reshape(B))
            obj_u.append(objective(xu, Qo, po).reshape(B))
    res = SolveResult(x, y, z, xv, H, C, torch.stack(pri_t), torch.stack(dual_t))
    if pri_u:
        res.pri_unscaled, res.dual_unscaled, res.obj_unscaled = (torch.stack(pri_u), torch.stack(dual_u),
                                                                 torch.stack(obj_u))
    return res


# --------------------------------------------------------------------------------------------
# Stage II exact ADMM (objective-gap oracle)
# --------------------------------------------------------------------------------------------
def exact_admm(Q, p, A0, zl, zu, rho_vec, sigma: float, iters: int, alpha: float = 1.6,
               state=None):
    """models/lu.py:13-47 iterated: factor K once, then x/z/y updates with relaxation alpha on both
    x and z.  Run in float64 to convergence it is the optimal-objective oracle (OSQP and Gurobi are
    not installed in this image)."""
    B, n, _ = Q.shape
    m = A0.shape[1]
    dt = Q.dtype
    if state is None:
        x = torch.zeros((B, n, 1), dtype=dt); y = torch.zeros((B, m, 1), dtype=dt); z = torch.zeros((B, m, 1), dtype=dt)
    else:
        x, y, z = state
    K, _ = kkt_system(Q, p, A0, x, y, z, rho_vec, sigma)
    LU, piv = torch.linalg.lu_factor(K)
    inv_rho = 1 / rho_vec
    for _ in range(iters):
        rhs = torch.cat((sigma * x - p, z - inv_rho * y), dim=1)
        xv = torch.linalg.lu_solve(LU, piv, rhs)
        xt, v = xv[:, :n, :], xv[:, n:, :]
        z_mid = z + inv_rho * (v - y)
        x = alpha * xt + (1 - alpha) * x
        z_rel = alpha * z_mid + (1 - alpha) * z
        z = torch.max(torch.min(z_rel + inv_rho * y, zu), zl)
        y = y + rho_vec * (z_rel - z)
    return x, y, z
