"""Pins the CPU oracle (oracle/iadmm_oracle.py) to outputs of the reference's own modules
(tests/golden/*.npz, produced by tests/golden/make_golden.py in the build container).

Tolerances: the oracle's dense form repeats the reference's arithmetic, so single steps agree to a
few fp32 ulps; the K-step trajectories agree far inside the reference's own fp32-vs-fp64 drift.
"""
import numpy as np
import pytest
import torch

from oracle import iadmm_oracle as orc
from helpers import load_golden, golden_params, golden_qp, rel_err, t


def test_state_dict_contract():
    g = load_golden("init_contract")
    prm = orc.lstm_parameters(8, 5)
    assert list(g["keys"]) == list(orc.PARAM_ORDER)
    for k in orc.PARAM_ORDER:
        assert tuple(g["shape_" + k]) == tuple(prm[k].shape), k


@pytest.mark.parametrize("name", ["step_small", "step_ineq_only", "step_eq_only"])
@pytest.mark.parametrize("form", ["dense", "block"])
def test_single_step(name, form):
    g = load_golden(name)
    B, n, mi, me, h, tt = (int(v) for v in g["meta"])
    prm, qp = golden_params(g), golden_qp(g)
    st = {k: t(g["in_" + k]) for k in ("x", "y", "z", "xv", "H", "C")}
    x, y, z, xv, H, C, rho_vec = orc.lstm_step(prm, tt, mi, me, st["x"], st["y"], st["z"], st["xv"], float(g["sigma"]),
                                               st["H"], st["C"], qp["Q"], qp["p"], qp["A0"], qp["zl"], qp["zu"], form=form)
    tol = 2e-6 if form == "dense" else 2e-5
    for k, v in dict(x=x, y=y, z=z, xv=xv, H=H, C=C, rho_vec=rho_vec).items():
        assert rel_err(v, g["out_" + k]) < tol, (k, rel_err(v, g["out_" + k]))
    K, rhs = orc.kkt_system(qp["Q"], qp["p"], qp["A0"], st["x"], st["y"], st["z"], rho_vec, float(g["sigma"]))
    assert torch.equal(K, t(g["out_K"]))
    assert torch.equal(rhs, t(g["out_rhs"]))


@pytest.mark.parametrize("name", ["ruiz_small", "ruiz_zero_rows", "ruiz_c1"])
def test_ruiz(name):
    g = load_golden(name)
    qp = golden_qp(g)
    Q, p, A0, zl, zu, sc = orc.ruiz_equilibrate(qp["Q"], qp["p"], qp["A0"], qp["zl"], qp["zu"], int(g["meta"][4]))
    for k, v in dict(Q=Q, p=p, A0=A0, zl=zl, zu=zu).items():
        assert rel_err(v, g["out_" + k]) < 1e-6, (k, rel_err(v, g["out_" + k]))
    assert rel_err(sc.d, g["out_d"]) < 1e-6 and rel_err(sc.e, g["out_e"]) < 1e-6
    assert rel_err(sc.c, g["out_c"]) < 1e-6 and rel_err(sc.cinv, g["out_cinv"]) < 1e-6
    assert rel_err(torch.reciprocal(sc.d), g["out_dinv"]) < 1e-6
    assert rel_err(torch.reciprocal(sc.e), g["out_einv"]) < 1e-6
    # -inf lower bounds of inequality rows survive the scaling
    assert np.array_equal(np.isinf(zl.numpy()), np.isinf(g["out_zl"]))


@pytest.mark.parametrize("name,tol", [("solve_small", 2e-5), ("solve_small_scaled", 2e-5), ("solve_small_bigw", 1e-4),
                                      ("solve_c1", 5e-5), ("solve_c1_scaled", 5e-5)])
@pytest.mark.parametrize("form", ["dense", "block"])
def test_solve_trajectory(name, tol, form):
    g = load_golden(name)
    B, n, mi, me, h, K, scaled = (int(v) for v in g["meta"])
    prm, qp = golden_params(g), golden_qp(g)
    Q, p, A0, zl, zu = (qp[k] for k in ("Q", "p", "A0", "zl", "zu"))
    sc, orig = None, None
    if scaled:
        orig = (Q, p, A0)
        Q, p, A0, zl, zu, sc = orc.ruiz_equilibrate(Q, p, A0, zl, zu, 10)
    r = orc.solve(prm, K, mi, me, Q, p, A0, zl, zu, float(g["sigma"]), h, scaling=sc, original=orig, form=form)
    for k in ("x", "y", "z", "xv"):
        e = rel_err(getattr(r, k), g["f32_" + k])
        assert e < tol, (k, e)
    assert rel_err(r.pri, g["f32_pri"]) < tol and rel_err(r.dual, g["f32_dual"]) < tol
    assert rel_err(r.H, g["f32_H"]) < tol and rel_err(r.C, g["f32_C"]) < tol
    if scaled:
        assert rel_err(r.pri_unscaled, g["f32_pri_u"]) < tol
        assert rel_err(r.dual_unscaled, g["f32_dual_u"]) < tol
        assert rel_err(r.obj_unscaled, g["f32_obj_u"]) < 10 * tol


def test_fp64_tiebreak():
    """The fp64 oracle reproduces the fp64 reference run to ~1e-12, so either can arbitrate."""
    g = load_golden("solve_small_scaled")
    B, n, mi, me, h, K, scaled = (int(v) for v in g["meta"])
    prm = golden_params(g, torch.float64)
    qp = golden_qp(g, torch.float64)
    Q, p, A0, zl, zu, sc = orc.ruiz_equilibrate(qp["Q"], qp["p"], qp["A0"], qp["zl"], qp["zu"], 10)
    r = orc.solve(prm, K, mi, me, Q, p, A0, zl, zu, float(g["sigma"]), h)
    for k in ("x", "y", "z", "xv"):
        assert rel_err(getattr(r, k), g["f64_" + k]) < 1e-9, k


def test_generator_feasible():
    """generate_data.py:72: c is built so that x = A^+ b is feasible."""
    qp = orc.qp_instances(2, 30, 10, 10, seed=1, dtype=torch.float64)
    x = torch.linalg.pinv(qp["A"]) @ qp["b"]
    assert float(orc.eq_violation(x, qp["A"], qp["b"]).max()) < 1e-4
    assert float(orc.ineq_violation(x, qp["G"], qp["c"]).max()) < 1e-4
    assert torch.isinf(qp["zl"][:, :10]).all() and torch.equal(qp["zl"][:, 10:], qp["zu"][:, 10:])


def test_exact_admm_converges():
    """Stage II (models/lu.py) in fp64 converges on a feasible instance: objective-gap oracle."""
    qp = orc.qp_instances(2, 20, 8, 8, seed=2, dtype=torch.float64)
    rho_vec = torch.full((2, 16, 1), 0.1, dtype=torch.float64)
    rho_vec[:, 8:] *= 1e3
    x, y, z = orc.exact_admm(qp["Q"], qp["p"], qp["A0"], qp["zl"], qp["zu"], rho_vec, 1e-6, 3000)
    pri, dual, _ = orc.primal_dual_residuals(x, y, z, qp["Q"], qp["p"], qp["A0"])
    assert float(pri.max()) < 1e-6 and float(dual.max()) < 1e-6
