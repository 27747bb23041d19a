"""Stage II linear algebra (csrc/lu.cu) against LAPACK semantics: `torch.lu` / `torch.lu_solve` as the reference
calls them in models/lu.py:30-35.  The factors are checked three ways: P K = L U reconstruction in fp64, the
pivot sequence against LAPACK's own (first maximum of every column), and solutions against an fp64 solve with
the library's fp32 error on the same system as the yardstick.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _kkt_like(B, n, m, seed):
    """Quasi-definite KKT matrices as LSTM.forward hands them to Stage II (models/lstm.py:61-62)."""
    g = torch.Generator().manual_seed(seed)
    M = torch.randn(B, n, n, generator=g) * (torch.rand(B, n, n, generator=g) < 0.15)
    Q = M @ M.transpose(1, 2) / n + 6e-6 * torch.eye(n)
    A = torch.randn(B, m, n, generator=g) * (torch.rand(B, m, n, generator=g) < 0.15)
    rho = torch.full((B, m), 0.5)
    rho[:, m // 2:] *= 1e3
    K = torch.cat((torch.cat((Q, A.transpose(1, 2)), 2), torch.cat((A, torch.diag_embed(-1 / rho)), 2)), 1)
    return K.to(DEV)


def _unpack(lu):
    N = lu.shape[-1]
    L = torch.tril(lu.double(), -1) + torch.eye(N, dtype=torch.float64, device=lu.device)
    return L, torch.triu(lu.double())


@pytest.mark.parametrize("B,N", [(3, 1), (4, 5), (3, 16), (3, 17), (2, 100), (2, 1030), (2, 2000), (1, 2100)])
def test_factor_reconstructs_permuted_matrix(B, N):
    import iadmm_b200.lu as lum
    g = torch.Generator().manual_seed(N)
    K = torch.randn(B, N, N, generator=g).to(DEV)
    lu, piv, info = lum.lu_factor(K)
    assert int(info.abs().max()) == 0
    L, U = _unpack(lu)
    perm = piv[1].long()
    assert all(sorted(perm[b].tolist()) == list(range(N)) for b in range(B))
    PK = torch.gather(K.double(), 1, perm[:, :, None].expand(B, N, N))
    err = (L @ U - PK).abs().max() / K.abs().max()
    assert float(err) < 2e-6 * max(1, N) ** 0.5, float(err)
    assert float(torch.tril(lu, -1).abs().max()) <= 1.0 + 1e-6          # partial pivoting: |L| <= 1


@pytest.mark.parametrize("N", [16, 64, 200])
def test_pivot_sequence_is_lapacks(N):
    """Same interchanges as getrf on matrices whose column maxima are well separated."""
    import iadmm_b200.lu as lum
    g = torch.Generator().manual_seed(7 + N)
    K = torch.randn(4, N, N, generator=g).to(DEV)
    _, piv, _ = lum.lu_factor(K)
    _, ref = torch.linalg.lu_factor(K.cpu().double())
    agree = (piv[0].cpu() == (ref - 1)).float().mean()
    assert float(agree) > 0.99, float(agree)


def test_first_maximum_on_ties():
    """isamax picks the FIRST row among equal magnitudes (also -v vs +v)."""
    import iadmm_b200.lu as lum
    K = torch.tensor([[[1., 2., 3.], [-1., 5., 1.], [1., 0., 7.]]], device=DEV)
    _, piv, _ = lum.lu_factor(K)
    assert piv[0, 0, 0].item() == 0


@pytest.mark.parametrize("B,n,m", [(3, 40, 26), (2, 500, 500), (2, 1000, 1000)])
def test_solve_on_kkt_systems(B, n, m):
    import iadmm_b200.lu as lum
    K = _kkt_like(B, n, m, seed=n)
    N = n + m
    g = torch.Generator().manual_seed(3)
    rhs = torch.randn(B, N, 1, generator=g).to(DEV)
    lu, piv, info = lum.lu_factor(K)
    assert int(info.abs().max()) == 0
    x = lum.lu_solve(lu, piv, rhs)
    x64 = torch.linalg.solve(K.double(), rhs.double())
    lib = torch.linalg.lu_solve(*torch.linalg.lu_factor(K), rhs)
    err = float((x.double() - x64).norm() / x64.norm())
    err_lib = float((lib.double() - x64).norm() / x64.norm())
    res = float((K.double() @ x.double() - rhs.double()).norm() / rhs.double().norm())
    assert res < 1e-4, res
    assert err < 5 * err_lib + 1e-6, (err, err_lib)


def test_singular_matrix_reports_info():
    import iadmm_b200.lu as lum
    K = torch.randn(2, 8, 8, device=DEV)
    K[1, :, 3] = 0.0
    _, _, info = lum.lu_factor(K)
    assert info[0].item() == 0 and info[1].item() == 4


def test_size_limit_is_an_error():
    import iadmm_b200 as ia
    import iadmm_b200.lu as lum
    with pytest.raises(ia.IadmmError):
        lum.lu_factor(torch.zeros(1, 4100, 4100, device=DEV))
