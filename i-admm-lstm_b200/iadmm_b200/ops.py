"""torch custom ops around the C ABI (north_star: "the model's forward path calls one thin C-ABI torch custom op").

    torch.ops.iadmm.solve       -> iadmm_solve        (K unrolled iterations + residual/metric traces, in place)
    torch.ops.iadmm.ruiz        -> iadmm_ruiz         (Scaling.scale_data)
    torch.ops.iadmm.residuals   -> iadmm_residuals    (primal_dual_loss)
    torch.ops.iadmm.build_kkt   -> iadmm_build_kkt    (A_tild / b_tild / rho_vec of LSTM.forward's return tuple)
    torch.ops.iadmm.kkt_penalty_diagonal -> iadmm_kkt_penalty_diagonal (the t-dependent diagonal of a kept A_tild)
    torch.ops.iadmm.sparse_pack / solve_sparse -> iadmm_sparse_pack / iadmm_solve_sparse (sparse problem families)

Each op is a few lines: it turns tensors into device pointers and calls the library on the current CUDA stream of the
tensors' device.  They are registered for the CUDA dispatch key ONLY -- calling one with CPU tensors fails in the
dispatcher ("no kernel for CPU"): there is no fallback.  State and trace tensors are declared as mutated arguments, so
the ops are safe under torch's functionalisation / graph capture; nothing is allocated inside (the caller passes the
workspace), nothing synchronises.
"""
from typing import Optional

import torch
from torch import Tensor

from . import _lib

_P = _lib.ptr


@torch.library.custom_op("iadmm::solve", mutates_args=("x", "y", "z", "xv", "H", "C", "pri", "dual", "pri_u", "dual_u",
                                                        "metrics", "workspace"), device_types="cuda")
def solve(packed: Tensor, Q: Tensor, p: Tensor, A0: Tensor, zl: Tensor, zu: Tensor,
          sd: Optional[Tensor], se: Optional[Tensor], sc: Optional[Tensor],
          x: Tensor, y: Tensor, z: Tensor, xv: Tensor, H: Tensor, C: Tensor,
          pri: Optional[Tensor], dual: Optional[Tensor], pri_u: Optional[Tensor], dual_u: Optional[Tensor],
          metrics: Optional[Tensor], workspace: Tensor,
          num_ineq: int, num_eq: int, h: int, length: int, t0: int, K: int, sigma: float, mode: int, flags: int) -> None:
    B, n = Q.shape[0], Q.shape[1]
    with torch.cuda.device(Q.device):
        _lib.check(_lib.lib().iadmm_solve(_P(packed), _P(Q), _P(p), _P(A0), _P(zl), _P(zu), _P(sd), _P(se), _P(sc),
                                          _P(x), _P(y), _P(z), _P(xv), _P(H), _P(C),
                                          _P(pri), _P(dual), _P(pri_u), _P(dual_u), _P(metrics),
                                          B, n, num_ineq, num_eq, h, length, t0, K, sigma, mode, flags,
                                          _P(workspace), workspace.numel(), _lib.stream_ptr()))


@torch.library.custom_op("iadmm::ruiz", mutates_args=("Qs", "ps", "A0s", "zls", "zus", "d", "e", "c", "workspace"),
                         device_types="cuda")
def ruiz(Q: Tensor, p: Tensor, A0: Tensor, zl: Tensor, zu: Tensor,
         Qs: Tensor, ps: Tensor, A0s: Tensor, zls: Tensor, zus: Tensor, d: Tensor, e: Tensor, c: Tensor,
         workspace: Tensor, iterations: int) -> None:
    B, n, m = Q.shape[0], Q.shape[1], A0.shape[1]
    with torch.cuda.device(Q.device):
        _lib.check(_lib.lib().iadmm_ruiz(_P(Q), _P(p), _P(A0), _P(zl), _P(zu), _P(Qs), _P(ps), _P(A0s), _P(zls), _P(zus),
                                         _P(d), _P(e), _P(c), B, n, m, iterations, _P(workspace), workspace.numel(),
                                         _lib.stream_ptr()))


@torch.library.custom_op("iadmm::residuals", mutates_args=("pri", "dual", "workspace"), device_types="cuda")
def residuals(x: Tensor, y: Tensor, z: Tensor, Q: Tensor, p: Tensor, A0: Tensor, pri: Tensor, dual: Tensor,
              workspace: Tensor) -> None:
    B, n, m = Q.shape[0], Q.shape[1], A0.shape[1]
    with torch.cuda.device(Q.device):
        _lib.check(_lib.lib().iadmm_residuals(_P(x), _P(y), _P(z), _P(Q), _P(p), _P(A0), _P(pri), _P(dual), B, n, m,
                                              _P(workspace), workspace.numel(), _lib.stream_ptr()))


@torch.library.custom_op("iadmm::build_kkt", mutates_args=("Kmat", "rhs", "rho_vec"), device_types="cuda")
def build_kkt(packed: Tensor, Q: Tensor, p: Tensor, A0: Tensor, x: Tensor, y: Tensor, z: Tensor,
              Kmat: Optional[Tensor], rhs: Tensor, rho_vec: Tensor,
              num_ineq: int, num_eq: int, h: int, length: int, t: int, sigma: float) -> None:
    B, n = Q.shape[0], Q.shape[1]
    with torch.cuda.device(Q.device):
        _lib.check(_lib.lib().iadmm_build_kkt(_P(packed), _P(Q), _P(p), _P(A0), _P(x), _P(y), _P(z), _P(Kmat), _P(rhs),
                                              _P(rho_vec), B, n, num_ineq, num_eq, h, length, t, sigma, _lib.stream_ptr()))


@torch.library.custom_op("iadmm::kkt_penalty_diagonal", mutates_args=("Kmat",), device_types="cuda")
def kkt_penalty_diagonal(packed: Tensor, Kmat: Tensor, n: int, num_ineq: int, num_eq: int, h: int, length: int, t: int) -> None:
    with torch.cuda.device(Kmat.device):
        _lib.check(_lib.lib().iadmm_kkt_penalty_diagonal(_P(packed), _P(Kmat), Kmat.shape[0], n, num_ineq, num_eq, h, length, t,
                                                         _lib.stream_ptr()))


@torch.library.custom_op("iadmm::sparse_pack", mutates_args=("packed", "nnz"), device_types="cuda")
def sparse_pack(M: Tensor, packed: Tensor, nnz: Tensor, cap: int) -> None:
    B, rows, n = M.shape
    with torch.cuda.device(M.device):
        _lib.check(_lib.lib().iadmm_sparse_pack(_P(M), B, rows, n, cap, _P(packed), packed.numel(), _P(nnz), _lib.stream_ptr()))


@torch.library.custom_op("iadmm::solve_sparse", mutates_args=("x", "y", "z", "xv", "H", "C", "pri", "dual", "pri_u", "dual_u",
                                                               "metrics", "workspace"), device_types="cuda")
def solve_sparse(packed: Tensor, Q: Optional[Tensor], Q_sp: Optional[Tensor], q_cap: int, Q_blk: Optional[Tensor], p: Tensor,
                 A0: Optional[Tensor], A0_sp: Optional[Tensor], a_cap: int, A0_blk: Optional[Tensor], zl: Tensor, zu: Tensor,
                 sd: Optional[Tensor], se: Optional[Tensor], sc: Optional[Tensor],
                 x: Tensor, y: Tensor, z: Tensor, xv: Tensor, H: Tensor, C: Tensor,
                 pri: Optional[Tensor], dual: Optional[Tensor], pri_u: Optional[Tensor], dual_u: Optional[Tensor],
                 metrics: Optional[Tensor], workspace: Tensor,
                 num_ineq: int, num_eq: int, h: int, length: int, t0: int, K: int, sigma: float, mode: int, flags: int) -> None:
    B, n = x.shape[0], x.shape[1]
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().iadmm_solve_sparse(_P(packed), _P(Q), _P(Q_sp), q_cap, _P(Q_blk), _P(p), _P(A0), _P(A0_sp), a_cap, _P(A0_blk),
                                                 _P(zl), _P(zu),
                                                 _P(sd), _P(se), _P(sc), _P(x), _P(y), _P(z), _P(xv), _P(H), _P(C),
                                                 _P(pri), _P(dual), _P(pri_u), _P(dual_u), _P(metrics),
                                                 B, n, num_ineq, num_eq, h, length, t0, K, sigma, mode, flags,
                                                 _P(workspace), workspace.numel(), _lib.stream_ptr()))


@torch.library.custom_op("iadmm::block_mask", mutates_args=("blocks", "nonempty"), device_types="cuda")
def block_mask(M: Tensor, blocks: Tensor, nonempty: Tensor) -> None:
    B, rows, n = M.shape
    with torch.cuda.device(M.device):
        _lib.check(_lib.lib().iadmm_block_mask(_P(M), B, rows, n, _P(blocks), blocks.numel(), _P(nonempty), _lib.stream_ptr()))
