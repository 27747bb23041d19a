"""Golden vectors for the dataset side of the path: the files of tests/golden/datasets/ (written by the reference's own
generate_data.py, see make_dataset_fixtures.py) loaded the way main.py loads them and pushed through the REFERENCE modules.

    python tests/golden/make_dataset_golden.py            # build container only (/root/reference)

Per family it restates main.py:236-302's loader literally (np.array over the per-instance lists, `.toarray()` for the csc
families, `Q * 2`, and the counts num_var / num_ineq / num_eq main.py derives from the G and A entries of the FILE), then
runs the unmodified `Scaling.scale_data` + K x `LSTM.forward` + `primal_dual_loss` with exactly those counts -- which for
Random_QP (G = [A0; -A0]) and SVM (G without the identity rows of A0) do NOT add up to the rows of A0 -- and stores inputs,
counts, weights and outputs in dataset_<family>.npz.
"""
import gzip
import os
import pickle
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from models.lstm import LSTM            # noqa: E402  (reference)
from methods.scaling import Scaling     # noqa: E402  (reference)
from utils import primal_dual_loss      # noqa: E402  (reference)

from oracle.iadmm_oracle import lstm_parameters   # noqa: E402  (weight generator only)

SIGMA, H, K, ITES = 6e-6, 8, 3, 10
FAMILIES = {"QP": ("QP_12_5_4", "qp_{}.gz"), "QP_RHS": ("QP_RHS_12_5_4", "qp_rhs_{}.gz"), "Random_QP": ("Random_QP_10_6", "random_qp_{}.gz"),
            "Equality_QP": ("Equality_QP_10_4", "equality_qp_{}.gz"), "SVM": ("SVM_10_4", "svm_{}.gz")}


def main_py_loader(prob_type, data_path, pattern, ids):
    """main.py:236-302, per-instance branches kept as written there."""
    Q, p, A0, zl, zu = [], [], [], [], []
    dense = prob_type in ("QP", "QP_RHS")
    for j in ids:
        with gzip.open(os.path.join(data_path, pattern.format(j)), "rb") as f:
            gz_dict = pickle.load(f)
        get = (lambda k: gz_dict[k]) if dense else (lambda k: gz_dict[k].toarray())
        Q.append(get("Q")); p.append(get("p"))
        num_var = get("Q").shape[1]
        try:
            num_ineq = get("G").shape[0]
        except KeyError:
            num_ineq = 0
        try:
            num_eq = get("A").shape[0]
        except KeyError:
            num_eq = 0
        A0.append(get("A0")); zl.append(get("zl")); zu.append(get("zu"))
    t = lambda a: torch.tensor(np.array(a), dtype=torch.float32)       # noqa: E731
    return t(Q) * 2, t(p), t(A0), t(zl), t(zu), num_var, num_ineq, num_eq


def main():
    for seed, (family, (sub, pattern)) in enumerate(FAMILIES.items()):
        Q, p, A0, zl, zu, num_var, num_ineq, num_eq = main_py_loader(family, os.path.join(HERE, "datasets", sub), pattern, [0, 1, 2])
        B, m = Q.shape[0], A0.shape[1]
        prm = lstm_parameters(H, K, 40 + seed, scale=8.0)
        model = LSTM(None, 2, H, K, "cpu")
        with torch.no_grad():
            for k, v in prm.items():
                getattr(model, k).copy_(v)
        scaling = Scaling(num_var, m, ITES, "cpu")
        Qs, ps, As, zls, zus = scaling.scale_data(Q, p, A0, zl, zu)
        x = torch.zeros((B, num_var, 1)); y = torch.zeros((B, m, 1)); z = torch.zeros((B, m, 1))
        xv = torch.zeros((B, num_var + m, 1)); Ht = torch.zeros((B, num_var + m, H)); Ct = torch.zeros((B, num_var + m, H))
        pri, dual = [], []
        with torch.no_grad():
            for t in range(K):
                x, y, z, xv, Ht, Ct, A_tild, b_tild, rho_vec = model(t, num_ineq, num_eq, x, y, z, xv, SIGMA, Ht, Ct, Q=Qs, p=ps, A0=As,
                                                                     lb=None, ub=None, zl=zls, zu=zus)
                pr, du, _ = primal_dual_loss(x, y, z, Qs, ps, As)
                pri.append(pr.reshape(B)); dual.append(du.reshape(B))
        out = dict(meta=np.array([B, num_var, num_ineq, num_eq, m, H, K, ITES]), sigma=np.float64(SIGMA),
                   in_Q=Q, in_p=p, in_A0=A0, in_zl=zl, in_zu=zu, sc_Q=Qs, sc_p=ps, sc_A0=As, sc_zl=zls, sc_zu=zus,
                   out_x=x, out_y=y, out_z=z, out_xv=xv, out_H=Ht, out_C=Ct, out_K=A_tild, out_rhs=b_tild, out_rho_vec=rho_vec,
                   out_pri=torch.stack(pri), out_dual=torch.stack(dual), **{"prm_" + k: v for k, v in prm.items()})
        np.savez_compressed(os.path.join(HERE, f"dataset_{family}.npz"),
                            **{k: (v.detach().numpy() if torch.is_tensor(v) else v) for k, v in out.items()})
        print(family, "B", B, "n", num_var, "file counts (num_ineq, num_eq)", (num_ineq, num_eq), "rows of A0", m,
              "rho classes", sorted(set(np.round(rho_vec.reshape(-1).numpy(), 4).tolist())))


if __name__ == "__main__":
    main()
