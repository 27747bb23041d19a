"""Stage-II kernels at config-2 size: factor + solve time vs the library (cuSOLVER/cuBLAS via torch.linalg)."""
import os, sys, json, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "i-admm-lstm_b200"))
import iadmm_b200.lu as lum

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

out = []
for B, N in [(8, 200), (64, 2000), (256, 2000)]:
    g = torch.Generator().manual_seed(1)
    K = torch.randn(B, N, N, generator=g).cuda()
    K = K + K.transpose(1, 2)
    rhs = torch.randn(B, N, 1, device="cuda")
    lu, piv, _ = lum.lu_factor(K)
    t_f = timed(lambda: lum.lu_factor(K), 2)
    t_s = timed(lambda: lum.lu_solve(lu, piv, rhs), 5)
    llu, lpiv = torch.linalg.lu_factor(K)
    t_lf = timed(lambda: torch.linalg.lu_factor(K), 2)
    t_ls = timed(lambda: torch.linalg.lu_solve(llu, lpiv, rhs), 5)
    x = lum.lu_solve(lu, piv, rhs); xl = torch.linalg.lu_solve(llu, lpiv, rhs)
    res = float((K @ x - rhs).norm() / rhs.norm()); resl = float((K @ xl - rhs).norm() / rhs.norm())
    out.append(dict(B=B, N=N, factor_ms=t_f, solve_ms=t_s, lib_factor_ms=t_lf, lib_solve_ms=t_ls, resid=res, lib_resid=resl))
    print(json.dumps(out[-1]), flush=True)
