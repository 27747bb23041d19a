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
