"""iadmm_b200: B200-native implementation of the I-ADMM-LSTM unrolled solve path.

Public surface = the reference's own call boundary for this path:
    LSTM            (models/lstm.py)        + LSTM.solve for K fused iterations
    Scaling         (methods/scaling.py)
    LU              (models/lu.py; library-backed Stage II, outside the accelerated path)
    primal_dual_loss, obj_fn, ineq_dist, eq_dist, lb_dist, ub_dist   (utils.py)
All compute goes through the C ABI of libiadmm_b200.so (include/iadmm.h); there is no fallback.
"""
from ._lib import IadmmError, LIB_PATH, GATE_MODES, lib
from .lstm import LSTM, SolveResult, SparseBatch
from .scaling import Scaling
from .lu import LU
from . import data
from .utils import primal_dual_loss, obj_fn, ineq_dist, eq_dist, lb_dist, ub_dist

__all__ = ["LSTM", "SolveResult", "SparseBatch", "Scaling", "LU", "primal_dual_loss", "obj_fn", "ineq_dist", "eq_dist",
           "lb_dist", "ub_dist", "data", "IadmmError", "LIB_PATH", "GATE_MODES", "lib"]
