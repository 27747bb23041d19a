"""Generate the golden fixtures in this directory by running the REFERENCE's own modules.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports ``models/lstm.py``, ``methods/scaling.py`` and ``utils.py`` from ``/root/reference``
unmodified, drives them exactly as ``main.py`` does (zero state main.py:837-843, loop :874-887,
residuals :346/:955, un-scaling :922,:923,:940) on seeded inputs, and stores inputs, weights and
outputs as ``*.npz``.  The tests never import the reference; they read these files.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from models.lstm import LSTM            # noqa: E402  (reference)
from methods.scaling import Scaling     # noqa: E402  (reference)
from utils import primal_dual_loss, obj_fn, ineq_dist, eq_dist   # noqa: E402  (reference)

from oracle.iadmm_oracle import qp_instances, lstm_parameters   # noqa: E402  (input generators only)

SIGMA = 6e-6   # configs/QP.yaml:14


def ref_model(prm, h, K, dtype):
    model = LSTM(None, 2, h, K, "cpu")
    with torch.no_grad():
        for k, v in prm.items():
            getattr(model, k).copy_(v)
    if dtype == torch.float64:
        model = model.double()
    return model.eval()


def npify(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def case_step(name, B, n, mi, me, h, seed):
    """One ``LSTM.forward`` from a random non-zero state, including the returned K, rhs, rho_vec."""
    m = mi + me
    qp = qp_instances(B, n, mi, me, seed)
    prm = lstm_parameters(h, 4, seed, scale=8.0)
    g = torch.Generator().manual_seed(seed + 1)
    st = dict(x=torch.randn((B, n, 1), generator=g), y=torch.randn((B, m, 1), generator=g),
              z=torch.randn((B, m, 1), generator=g), xv=torch.randn((B, n + m, 1), generator=g),
              H=torch.tanh(torch.randn((B, n + m, h), generator=g)), C=torch.randn((B, n + m, h), generator=g))
    model = ref_model(prm, h, 4, torch.float32)
    t = 2
    with torch.no_grad():
        out = model(t, mi, me, st["x"], st["y"], st["z"], st["xv"], SIGMA, st["H"], st["C"],
                    Q=qp["Q"], p=qp["p"], A0=qp["A0"], lb=None, ub=None, zl=qp["zl"], zu=qp["zu"])
    keys = ("x", "y", "z", "xv", "H", "C", "K", "rhs", "rho_vec")
    blob = dict(meta=np.array([B, n, mi, me, h, t]), sigma=SIGMA)
    blob.update({f"in_{k}": v for k, v in npify(st).items()})
    blob.update({f"qp_{k}": v for k, v in npify(qp).items() if k in ("Q", "p", "A0", "zl", "zu")})
    blob.update({f"prm_{k}": v for k, v in npify(prm).items()})
    blob.update({f"out_{k}": v.detach().numpy() for k, v in zip(keys, out)})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **blob)
    print("wrote", name)


def case_ruiz(name, B, n, mi, me, seed, ites=10, zero_row=False):
    qp = qp_instances(B, n, mi, me, seed)
    if zero_row:   # exercise the "norm clamps to 1e-4 -> 1.0" branch (scaling.py:37)
        qp["A0"][0, 1, :] = 0.0
        qp["A0"][:, :, 2] = 0.0
        qp["Q"][:, 2, 2] = 0.0
    sc = Scaling(n, mi + me, ites, "cpu")
    Qs, ps, As, zls, zus = sc.scale_data(qp["Q"], qp["p"], qp["A0"], qp["zl"], qp["zu"])
    blob = dict(meta=np.array([B, n, mi, me, ites]))
    blob.update({f"qp_{k}": v for k, v in npify(qp).items() if k in ("Q", "p", "A0", "zl", "zu")})
    blob.update(out_Q=Qs.numpy(), out_p=ps.numpy(), out_A0=As.numpy(), out_zl=zls.numpy(), out_zu=zus.numpy(),
                out_d=sc.D.diagonal(dim1=1, dim2=2).numpy(), out_e=sc.E.diagonal(dim1=1, dim2=2).numpy(),
                out_dinv=sc.D_inv.diagonal(dim1=1, dim2=2).numpy(),
                out_einv=sc.Einv.diagonal(dim1=1, dim2=2).numpy(),
                out_c=sc.c.numpy(), out_cinv=sc.cinv.numpy())
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **blob)
    print("wrote", name)


def case_solve(name, B, n, mi, me, h, K, seed, scaling, wscale=1.0, store_inputs=True):
    """K-iteration solve from the zero state, as main.py's test loop runs it (fp32), plus the same
    solve in fp64 (tie-breaker).  Traces: residuals on the solve's own data (main.py:346) and, with
    scaling, on the un-scaled iterates and original data (main.py:922-955)."""
    m = mi + me
    qp = qp_instances(B, n, mi, me, seed)
    prm = lstm_parameters(h, K, seed, scale=wscale)
    blob = dict(meta=np.array([B, n, mi, me, h, K, int(scaling)]), sigma=SIGMA, wscale=wscale, seed=seed)
    if store_inputs:
        blob.update({f"qp_{k}": v for k, v in npify(qp).items() if k in ("Q", "p", "A0", "zl", "zu")})
        blob.update({f"prm_{k}": v for k, v in npify(prm).items()})
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        torch.set_default_dtype(dt)     # the reference allocates its helpers in the default dtype
        model = ref_model({k: v.to(dt) for k, v in prm.items()}, h, K, dt)
        Q, p, A0, zl, zu = (qp[k].to(dt) for k in ("Q", "p", "A0", "zl", "zu"))
        Q0, p0, A00 = Q, p, A0
        sc = None
        if scaling:
            sc = Scaling(n, m, 10, "cpu")
            Q, p, A0, zl, zu = sc.scale_data(Q, p, A0, zl, zu)
        x = torch.zeros((B, n, 1), dtype=dt); y = torch.zeros((B, m, 1), dtype=dt)
        z = torch.zeros((B, m, 1), dtype=dt); xv = torch.zeros((B, n + m, 1), dtype=dt)
        H = torch.zeros((B, n + m, h), dtype=dt); C = torch.zeros((B, n + m, h), dtype=dt)
        pri, dual, pri_u, dual_u, obj_u, ls = [], [], [], [], [], []
        vio = []     # [K, 4, B]: ineq max, ineq mean, eq max, eq mean on the original data (main.py:959-968)
        G0, c0 = qp["G"].to(dt), qp["c"].to(dt)
        Aeq0, b0 = qp["A"].to(dt), qp["b"].to(dt)
        with torch.no_grad():
            for t in range(K):
                x, y, z, xv, H, C, Kmat, rhs, rho_vec = model(t, mi, me, x, y, z, xv, SIGMA, H, C, Q=Q, p=p, A0=A0,
                                                              lb=None, ub=None, zl=zl, zu=zu)
                pr, du, _ = primal_dual_loss(x, y, z, Q, p, A0)
                pri.append(pr.reshape(B).numpy()); dual.append(du.reshape(B).numpy())
                ls.append(torch.linalg.vector_norm(torch.bmm(Kmat, xv) - rhs, dim=(1, 2)).numpy())   # main.py:952
                if scaling:
                    xu = torch.bmm(sc.D, x); zu_ = torch.bmm(sc.Einv, z); yu = torch.bmm(sc.cinv * sc.E, y)
                    pr, du, _ = primal_dual_loss(xu, yu, zu_, Q0, p0, A00)
                    pri_u.append(pr.reshape(B).numpy()); dual_u.append(du.reshape(B).numpy())
                    obj_u.append(obj_fn(xu, Q=Q0, p=p0).reshape(B).numpy())
                    iv, ev = ineq_dist(xu, G=G0, c=c0), eq_dist(xu, A=Aeq0, b=b0)
                    vio.append(np.stack([iv.max(dim=1).values.reshape(B).numpy(), iv.mean(dim=1).reshape(B).numpy(),
                                         ev.max(dim=1).values.reshape(B).numpy(), ev.mean(dim=1).reshape(B).numpy()]))
        blob.update({f"{tag}_x": x.numpy(), f"{tag}_y": y.numpy(), f"{tag}_z": z.numpy(), f"{tag}_xv": xv.numpy(),
                     f"{tag}_pri": np.stack(pri), f"{tag}_dual": np.stack(dual), f"{tag}_ls": np.stack(ls)})
        if tag == "f32":
            blob.update(f32_H=H.numpy().astype(np.float32), f32_C=C.numpy().astype(np.float32))
        if scaling:
            blob.update({f"{tag}_pri_u": np.stack(pri_u), f"{tag}_dual_u": np.stack(dual_u), f"{tag}_obj_u": np.stack(obj_u),
                         f"{tag}_vio_u": np.stack(vio)})
    torch.set_default_dtype(torch.float32)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **blob)
    print("wrote", name)


def case_init(name, h, K, seed):
    """The reference constructor's own parameter names/shapes under torch.manual_seed (state_dict contract)."""
    torch.manual_seed(seed)
    model = LSTM(None, 2, h, K, "cpu")
    sd = model.state_dict()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), keys=np.array(list(sd.keys())),
                        **{"shape_" + k: np.array(v.shape) for k, v in sd.items()})
    print("wrote", name)


if __name__ == "__main__":
    torch.set_num_threads(8)
    case_init("init_contract", 8, 5, 17)
    case_step("step_small", B=3, n=12, mi=5, me=7, h=8, seed=3)
    case_step("step_ineq_only", B=2, n=10, mi=6, me=0, h=8, seed=4)
    case_step("step_eq_only", B=2, n=10, mi=0, me=6, h=8, seed=5)
    case_ruiz("ruiz_small", B=3, n=12, mi=5, me=7, seed=6)
    case_ruiz("ruiz_zero_rows", B=2, n=12, mi=5, me=7, seed=7, zero_row=True)
    case_ruiz("ruiz_c1", B=2, n=100, mi=50, me=50, seed=8)
    case_solve("solve_small", B=3, n=12, mi=5, me=7, h=8, K=15, seed=9, scaling=False)
    case_solve("solve_small_scaled", B=3, n=12, mi=5, me=7, h=8, K=15, seed=10, scaling=True)
    case_solve("solve_small_bigw", B=3, n=12, mi=5, me=7, h=8, K=15, seed=11, scaling=True, wscale=10.0)
    case_solve("solve_c1", B=4, n=100, mi=50, me=50, h=64, K=100, seed=17, scaling=False)
    case_solve("solve_c1_scaled", B=4, n=100, mi=50, me=50, h=64, K=100, seed=18, scaling=True)
