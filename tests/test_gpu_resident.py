"""The on-chip-resident solve variant (csrc/resident.cu: one persistent CTA per instance, n+m <= 256, hidden 64)
against the streaming variant of the same library and against the CPU oracle.  Both variants implement
models/lstm.py:47-96 + utils.py:68-71; they differ in summation order only, so they agree to fp32 rounding noise.
"""
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _setup(B, n, mi, me, K, mode, seed=5, scale=True):
    import iadmm_b200 as ia
    from oracle import iadmm_oracle as orc
    qp = {k: v.to(DEV) for k, v in orc.qp_instances(B, n, mi, me, seed=seed).items()}
    prm = orc.lstm_parameters(64, K, seed=seed)
    model = ia.LSTM(None, 2, 64, K, DEV, gate_mode=mode)
    with torch.no_grad():
        for k, v in prm.items():
            getattr(model, k).copy_(v.to(DEV))
    sc = None
    data = (qp["Q"], qp["p"], qp["A0"], qp["zl"], qp["zu"])
    if scale:
        sc = ia.Scaling(n, mi + me, 10, DEV)
        data = sc.scale_data(*data)
    return model.eval(), data, sc, qp, prm


def _close(a, b, tol, what):
    assert rel_err(a, b) < tol, (what, rel_err(a, b))


@pytest.mark.parametrize("mode", ["tc_f16f8", "tc_3xfp16", "tc_1xfp16"])
@pytest.mark.parametrize("shape", [(5, 100, 50, 50), (3, 40, 12, 14), (2, 150, 60, 40), (3, 128, 64, 64), (2, 30, 0, 0), (160, 20, 5, 6)])
def test_resident_matches_streaming(shape, mode):
    B, n, mi, me = shape
    K = 30
    model, data, sc, _, _ = _setup(B, n, mi, me, K, mode)
    with torch.no_grad():
        a = model.solve(K, mi, me, *data, sigma=6e-6, scaling=sc, traces=True)
        s = model.solve(K, mi, me, *data, sigma=6e-6, scaling=sc, traces=True, streaming=True)
    tol = 2e-3 if mode == "tc_1xfp16" else 5e-5
    for name in ("x", "z", "xv", "H", "C"):
        _close(getattr(a, name), getattr(s, name), tol, name)
    if mi + me:
        _close(a.y, s.y, 20 * tol, "y")          # rho * ulp(z) noise on equality rows, see test_gpu_parity.py
    for name in ("pri", "dual", "pri_unscaled", "dual_unscaled"):
        _close(getattr(a, name), getattr(s, name), 10 * tol, name)
    assert torch.allclose(a.metrics, s.metrics, rtol=20 * tol, atol=1e-5)


def test_resident_vs_oracle_k100():
    """BASELINE config 1 shape (n=100, 50+50, h=64, K=100) against the CPU oracle at north_star's tolerance."""
    from oracle import iadmm_oracle as orc
    B, n, mi, me, K = 4, 100, 50, 50, 100
    model, data, sc, qp, prm = _setup(B, n, mi, me, K, "tc_f16f8", seed=11, scale=False)
    with torch.no_grad():
        r = model.solve(K, mi, me, *data, sigma=6e-6, traces=True)
    ref = orc.solve(prm, K, mi, me, *(qp[k].cpu() for k in ("Q", "p", "A0", "zl", "zu")), 6e-6, 64, form="block")
    ref64 = orc.solve({k: v.double() for k, v in prm.items()}, K, mi, me, *(qp[k].cpu().double() for k in ("Q", "p", "A0", "zl", "zu")),
                      6e-6, 64, form="block")
    for name in ("x", "z"):
        _close(getattr(r, name), getattr(ref, name), 1e-4, name)
    # y: 1e-4 against the fp32 oracle, or as close to the float64 result as the fp32 oracle itself (un-scaled data, seed 11:
    # the reference's own fp32 run is a few 1e-4 off its float64 run on y -- helpers.assert_parity states that explicitly)
    from helpers import assert_parity
    print("resident K=100 y", assert_parity("y", r.y, ref.y, ref64.y))
    _close(r.pri, ref.pri, 1e-4, "pri")
    _close(r.dual, ref.dual, 1e-4, "dual")


def test_resident_continuation_equals_one_shot():
    """Two calls (t0 = 0 and t0 = 12, state handed over) == one call of 30 iterations, bit for bit: the state that
    leaves the kernel (fp32 H, C, x, y, z, xv) is all the state there is."""
    B, n, mi, me, K = 3, 60, 20, 25, 30
    model, data, sc, _, _ = _setup(B, n, mi, me, K, "tc_f16f8")
    with torch.no_grad():
        one = model.solve(K, mi, me, *data, sigma=6e-6, traces=False)
        a = model.solve(12, mi, me, *data, sigma=6e-6, traces=False)
        b = model.solve(K - 12, mi, me, *data, sigma=6e-6, traces=False, t0=12, state=(a.x, a.y, a.z, a.xv, a.H, a.C))
    for name in ("x", "y", "z", "xv", "H", "C"):
        assert torch.equal(getattr(one, name), getattr(b, name)), name


def test_resident_is_batch_independent():
    """Instances are solved by independent CTAs: any sub-batch gives bit-identical results."""
    B, n, mi, me, K = 7, 50, 20, 20, 10
    model, data, sc, _, _ = _setup(B, n, mi, me, K, "tc_f16f8", scale=False)
    with torch.no_grad():
        full = model.solve(K, mi, me, *data, sigma=6e-6, traces=True)
        part = model.solve(K, mi, me, *(t[2:5].contiguous() for t in data), sigma=6e-6, traces=True)
    for name in ("x", "y", "z", "xv", "H", "C"):
        assert torch.equal(getattr(full, name)[2:5], getattr(part, name)), name
    assert torch.equal(full.pri[:, 2:5], part.pri)
