"""CPU tests of the dataset I/O on either side of the path: the reference's per-instance
gzip+pickle format, dense (QP) and sparse (Random_QP ...) families, the Q*2-on-load convention."""
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "i-admm-lstm_b200"))


def test_dense_family_roundtrip(tmp_path):
    from iadmm_b200 import data
    from oracle import iadmm_oracle as orc
    n, mi, me = 12, 5, 4
    qp = orc.qp_instances(3, n, mi, me, seed=4)
    d = data.dataset_dir(str(tmp_path), "QP", n, mi, me)
    assert d.endswith(f"QP_{n}_{mi}_{me}")
    for i in range(3):      # what generate_data.py:88-92 stores: Q0 = Q/2, column vectors, A0 = [G; A]
        inst = dict(Q=(qp["Q"][i] / 2).numpy(), p=qp["p"][i].numpy(), G=qp["G"][i].numpy(), c=qp["c"][i].numpy(),
                    A=qp["A"][i].numpy(), b=qp["b"][i].numpy(), A0=qp["A0"][i].numpy(), zl=qp["zl"][i].numpy(),
                    zu=qp["zu"][i].numpy(), x=np.zeros(n), y=np.zeros(mi + me))
        data.write_instance(data.instance_path(d, "QP", 10 + i), inst)
    assert os.path.basename(data.instance_path(d, "QP", 11)) == "qp_11.gz"
    batch, sizes = data.load_batch(d, "QP", [10, 11, 12], "cpu")
    assert sizes == dict(num_var=n, num_ineq=mi, num_eq=me)
    for k in ("Q", "p", "A0", "zl", "zu", "G", "c", "A", "b"):
        assert batch[k].dtype == torch.float32 and torch.equal(batch[k], qp[k]), k      # Q doubled back on load
    assert batch["p"].shape == (3, n, 1) and batch["zl"].shape == (3, mi + me, 1)


def test_sparse_family_is_densified(tmp_path):
    from iadmm_b200 import data
    rng = np.random.default_rng(0)
    n, mi = 10, 6
    d = data.dataset_dir(str(tmp_path), "Random_QP", n, mi)
    dense = []
    for i in range(2):      # generate_data.py:96-134 stores csc matrices for this family
        M = rng.standard_normal((n, n)) * (rng.random((n, n)) < 0.6)
        Q = M @ M.T
        G = rng.standard_normal((mi, n)) * (rng.random((mi, n)) < 0.6)
        inst = dict(Q=sp.csc_matrix(Q), p=sp.csc_matrix(rng.standard_normal((n, 1))), G=sp.csc_matrix(G),
                    c=sp.csc_matrix(rng.random((mi, 1))), A0=sp.csc_matrix(G), zl=sp.csc_matrix(-np.ones((mi, 1))),
                    zu=sp.csc_matrix(rng.random((mi, 1))))
        dense.append({k: v.toarray() for k, v in inst.items()})
        data.write_instance(data.instance_path(d, "Random_QP", i), inst)
    batch, sizes = data.load_batch(d, "Random_QP", [0, 1], "cpu")
    assert sizes == dict(num_var=n, num_ineq=mi, num_eq=0)
    assert np.allclose(batch["Q"].numpy(), 2 * np.stack([x["Q"] for x in dense]).astype(np.float32))
    assert np.allclose(batch["A0"].numpy(), np.stack([x["A0"] for x in dense]).astype(np.float32))
    assert batch["p"].shape == (2, n, 1)


def test_device_generator_matches_family():
    """generate_qp_batch (CPU device here) produces the QP family: diagonal Q in [0,1), feasible at x = A^+ b."""
    from iadmm_b200 import data
    d = data.generate_qp_batch(4, 30, 10, 10, seed=1, device="cpu")
    Q = d["Q"]
    assert torch.equal(Q, torch.diag_embed(Q.diagonal(dim1=1, dim2=2))) and 0 <= float(Q.min()) and float(Q.max()) < 1
    x = torch.linalg.pinv(d["A"].double()) @ d["b"].double()
    assert float((d["A"].double() @ x - d["b"].double()).abs().max()) < 1e-5
    assert float((d["G"].double() @ x - d["c"].double()).max()) < 1e-5
    assert torch.isinf(d["zl"][:, :10]).all() and torch.equal(d["zl"][:, 10:], d["zu"][:, 10:])
    s = data.generate_qp_batch(4, 30, 10, 10, seed=1, device="cpu", as_stored=True)
    assert torch.equal(2 * s["Q"], Q)


# ---- files written by the reference's own generate_data.py (tests/golden/datasets, make_dataset_fixtures.py) -------------
FAMILY_DIRS = {"QP": "QP_12_5_4", "QP_RHS": "QP_RHS_12_5_4", "Random_QP": "Random_QP_10_6", "Equality_QP": "Equality_QP_10_4",
               "SVM": "SVM_10_4"}
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _golden(family):
    z = np.load(os.path.join(GOLDEN, f"dataset_{family}.npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


import pytest   # noqa: E402


@pytest.mark.parametrize("family", list(FAMILY_DIRS))
def test_reads_files_written_by_generate_data(family):
    """`load_batch` on the reference writer's files == main.py:236-302's loader (restated in make_dataset_golden.py): the
    tensors, and the counts main.py derives from the G / A entries of the file."""
    from iadmm_b200 import data
    g = _golden(family)
    B, n, num_ineq, num_eq, m = (int(v) for v in g["meta"][:5])
    d = os.path.join(GOLDEN, "datasets", FAMILY_DIRS[family])
    assert os.path.basename(data.dataset_dir("x", family, *{"QP": (12, 5, 4), "QP_RHS": (12, 5, 4), "Random_QP": (10, 6, None),
                                                            "Equality_QP": (10, None, 4), "SVM": (10, 4, None)}[family])) == FAMILY_DIRS[family]
    batch, sizes = data.load_batch(d, family, [0, 1, 2], "cpu")
    assert sizes == dict(num_var=n, num_ineq=num_ineq, num_eq=num_eq)
    for k in ("Q", "p", "A0", "zl", "zu"):
        assert batch[k].dtype == torch.float32 and np.array_equal(batch[k].numpy(), g["in_" + k]), k
    assert batch["A0"].shape[1] == m


@pytest.mark.parametrize("family", list(FAMILY_DIRS))
def test_oracle_on_dataset_files_with_main_py_counts(family):
    """The oracle with the counts main.py passes (which need not add up to the rows of A0: Random_QP 12 + 0 vs 6 rows, SVM
    4 + 0 vs 14 rows) against the reference's Scaling + K x forward + primal_dual_loss on the same files."""
    from oracle import iadmm_oracle as orc
    from helpers import rel_err, t, golden_params
    from iadmm_b200.lstm import row_classes
    g = _golden(family)
    B, n, num_ineq, num_eq, m, h, K, ites = (int(v) for v in g["meta"])
    prm = golden_params(g)
    Q, p, A0, zl, zu, _ = orc.ruiz_equilibrate(*(t(g["in_" + k]) for k in ("Q", "p", "A0", "zl", "zu")), ites)
    for k, v in dict(Q=Q, p=p, A0=A0).items():
        assert rel_err(v, g["sc_" + k]) < 1e-6, k
    r = orc.solve(prm, K, num_ineq, num_eq, t(g["sc_Q"]), t(g["sc_p"]), t(g["sc_A0"]), t(g["sc_zl"]), t(g["sc_zu"]), float(g["sigma"]), h)
    for k in ("x", "y", "z", "xv", "H", "C", "pri", "dual"):
        assert rel_err(getattr(r, k), g["out_" + k]) < 2e-5, (k, rel_err(getattr(r, k), g["out_" + k]))
    # the kernels' (inequality rows, equality rows) for these counts reproduce the reference's rho_vec classes
    ni, ne = row_classes(num_ineq, num_eq, m)
    assert ni + ne == m
    rho = g["out_rho_vec"][0, :, 0]
    assert np.all(rho[:ni] == rho.min()) and (ne == 0 or np.all(rho[ni:] == rho.max()))
    rv, _ = orc.penalty_schedule(prm, K - 1, ni, ne, B)
    assert rel_err(rv, g["out_rho_vec"]) < 1e-6


def test_row_classes():
    from iadmm_b200.lstm import row_classes
    assert row_classes(5, 4, 9) == (5, 4)          # QP
    assert row_classes(12, 0, 6) == (6, 0)         # Random_QP: G = [A0; -A0]
    assert row_classes(4, 0, 14) == (14, 0)        # SVM: G without the identity rows
    assert row_classes(0, 4, 4) == (0, 4)          # Equality_QP
    assert row_classes(3, 9, 7) == (3, 4)          # the slice is clipped by the tensor like in torch
    assert row_classes(0, 0, 5) == (5, 0)
    with pytest.raises(ValueError):
        row_classes(2, 3, 9)                       # equality rows followed by inequality-class rows
    with pytest.raises(ValueError):
        row_classes(-1, 3, 9)
