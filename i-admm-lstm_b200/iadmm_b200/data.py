"""Data on either side of the solve path: the reference's on-disk instance format and a device-side generator.

On-disk format (written by generate_data.py:77-94, read by main.py:198-302 / :384-499 / :621-739): one file per
instance, `gzip(pickle(dict))` with the keys `Q, p, A0, zl, zu` (+ `G, c, A, b`, optional `lb, ub`, OSQP labels
`x, y`); dense numpy arrays for the `QP`/`QP_RHS` families, scipy sparse matrices (densified with `.toarray()` on
load) for the others.  main.py doubles `Q` on load (`*2`, :298/:467/:718) because generate_data.py stores
`0.5*diag` and hands OSQP `P = 2*Q0`.  File names follow main.py:202-232, directory names main.py:78-166.

`load_batch` reads a list of instance ids into pinned host staging buffers and copies them to the device
asynchronously, returning the tensors main.py builds (fp32, `[B, ., .]`).  `generate_qp_batch` restates the `QP`
family of generate_data.py:67-76 on the device (used by bench.py; the OSQP "solved" filter is skipped).
"""
import gzip
import os
import pickle

import numpy as np
import torch

FILE_PATTERN = {
    "QP_RHS": "qp_rhs_{id}.gz", "QP": "qp_{id}.gz", "Random_QP": "random_qp_{id}.gz",
    "Equality_QP": "equality_qp_{id}.gz", "SVM": "svm_{id}.gz", "QPLIB": "qplib_{qplib}_{id}.gz",
    "MM_MOSARQP2": "mosarqp2_{id}.gz", "MM_QSCSD6": "qscsd6_{id}.gz", "MM_QSCRS8": "qscrs8_{id}.gz",
    "MM_Q25FV47": "q25fv47_{id}.gz", "MM_QSHIP04L": "qship04l_{id}.gz", "MM_QSHIP08S": "qship08s_{id}.gz",
    "MM_CVXQP1_M": "cvxqp1_m_{id}.gz", "MM_CVXQP3_M": "cvxqp3_m_{id}.gz",
}


def dataset_dir(root, prob_type, num_var=None, num_ineq=None, num_eq=None, qplib_num=None):
    """Directory of a dataset below `root` (main.py:78-166)."""
    if prob_type in ("QP", "QP_RHS"):
        name = f"{prob_type}_{num_var}_{num_ineq}_{num_eq}"
    elif prob_type in ("Random_QP", "SVM"):
        name = f"{prob_type}_{num_var}_{num_ineq}"
    elif prob_type == "Equality_QP":
        name = f"{prob_type}_{num_var}_{num_eq}"
    elif prob_type == "QPLIB":
        name = f"{prob_type}_{qplib_num}"
    else:
        name = prob_type
    return os.path.join(root, name)


def instance_path(data_path, prob_type, instance_id, qplib_num=None):
    return os.path.join(data_path, FILE_PATTERN[prob_type].format(id=instance_id, qplib=qplib_num))


def _dense(a):
    return np.asarray(a.toarray() if hasattr(a, "toarray") else a)


def load_instance(path):
    """One instance as a dict of dense numpy arrays (sparse families are densified like main.py:243-296)."""
    with gzip.open(path, "rb") as f:
        raw = pickle.load(f)
    return {k: _dense(v) for k, v in raw.items()}


def write_instance(path, instance):
    """Write one instance in the reference's format (generate_data.py:88-92)."""
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with gzip.open(path, "wb") as f:
        pickle.dump({k: (v.detach().cpu().numpy() if torch.is_tensor(v) else v) for k, v in instance.items()}, f)


def load_batch(data_path, prob_type, ids, device, qplib_num=None, pin=True):
    """Instances `ids` as device tensors: Q [B,n,n] (doubled, main.py:298), p [B,n,1], A0 [B,m,n], zl, zu [B,m,1] and,
    when present, G, c, A, b, lb, ub.  Also returns the sizes main.py derives (num_var, num_ineq, num_eq): the row counts of the
    file's G and A entries (main.py:248-272), to be passed to `LSTM.forward` / `solve` unchanged -- they need not add up to the
    rows of A0 (Random_QP: G = [A0; -A0]; SVM: G without the identity rows), see `iadmm_b200.lstm.row_classes`."""
    insts = [load_instance(instance_path(data_path, prob_type, i, qplib_num)) for i in ids]
    out = {}
    dev = torch.device(device)
    use_pin = pin and dev.type == "cuda"
    for key in ("Q", "p", "A0", "zl", "zu", "G", "c", "A", "b", "lb", "ub"):
        if not all(key in d for d in insts):
            continue
        arr = np.stack([d[key] for d in insts]).astype(np.float32)
        if arr.ndim == 2:
            arr = arr[..., None]
        host = torch.from_numpy(arr)
        if use_pin:
            host = host.pin_memory()
        out[key] = host.to(dev, non_blocking=use_pin)
    out["Q"] = out["Q"] * 2
    sizes = dict(num_var=out["Q"].shape[1],
                 num_ineq=out["G"].shape[1] if "G" in out else 0,
                 num_eq=out["A"].shape[1] if "A" in out else 0)
    return out, sizes


def generate_qp_batch(batch, n, num_ineq, num_eq, seed, device, as_stored=False):
    """The `QP` family of generate_data.py:67-76 on `device`: Q0 = 0.5 diag(U[0,1)), p ~ U[0,1), A ~ N(0,1),
    b ~ U[-1,1), G ~ N(0,1), c = sum_j |G A^+| (feasible at x = A^+ b), A0 = [G; A], zl = [-inf; b], zu = [c; b].
    Returns the tensors as main.py sees them after loading (Q doubled) unless `as_stored`."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    Q0 = 0.5 * torch.diag_embed(torch.rand((batch, n), device=dev, generator=g))
    p = torch.rand((batch, n, 1), device=dev, generator=g)
    A = torch.randn((batch, num_eq, n), device=dev, generator=g)
    b = 2 * torch.rand((batch, num_eq, 1), device=dev, generator=g) - 1
    G = torch.randn((batch, num_ineq, n), device=dev, generator=g)
    c = torch.empty((batch, num_ineq, 1), device=dev)
    for lo in range(0, batch, 32):     # A^+ = A^T (A A^T)^-1 for the full-row-rank A of this family
        Ab, Gb = A[lo:lo + 32].double(), G[lo:lo + 32].double()
        c[lo:lo + 32] = torch.linalg.solve(Ab @ Ab.mT, (Gb @ Ab.mT).mT).mT.abs().sum(dim=2, keepdim=True).float()
    A0 = torch.cat((G, A), dim=1).contiguous()
    zl = torch.cat((torch.full_like(c, float("-inf")), b), dim=1).contiguous()
    zu = torch.cat((c, b), dim=1).contiguous()
    Q = Q0 if as_stored else 2 * Q0
    return dict(Q=Q, p=p, A0=A0, zl=zl, zu=zu, G=G, c=c, A=A, b=b)


def generate_family_batch(family, batch, num_var, num_ineq=0, num_eq=0, seed=0, device="cuda", density=None):
    """The sparse problem families of generate_data.py on `device`, as main.py sees them after loading (densified, Q doubled):
      Random_QP   (:96-134)  M = N(0,1) masked at 60 %, Q0 = (M M^T + 0.01 I)/2, A0 = N(0,1) masked at 60 % [num_ineq, n],
                             zl = -U[0,1), zu = U[0,1)
      Equality_QP (:136-175) same Q at 50 %, A0 = A = N(0,1) masked at 50 % [num_eq, n], zl = zu = b ~ N(0,1)
      SVM         (:177-228) n = num_var + num_ineq variables, Q0 = diag(I_num_var, 0), p = [0; lambda 1],
                             A0 = [[diag(b^) A^ , -I], [I_n]] with A^ masked at 50 %, zl = [-inf; -inf; 0], zu = [-1; +inf]
    `density` overrides the family's mask probability (e.g. 0.01 for a QPLIB-like truly sparse instance).  The OSQP "solved"
    filter of the reference is skipped.  Returns Q, p, A0, zl, zu (+ the counts main.py derives)."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    B = batch

    def randn(*s):
        return torch.randn(s, device=dev, generator=g)

    def rand(*s):
        return torch.rand(s, device=dev, generator=g)

    if family in ("Random_QP", "Equality_QP"):
        n = num_var
        sp_ = density if density is not None else (0.6 if family == "Random_QP" else 0.5)
        M = randn(B, n, n) * (rand(B, n, n) < sp_)
        Q0 = (M @ M.mT + 0.01 * torch.eye(n, device=dev)) * 0.5
        p = randn(B, n, 1)
        if family == "Random_QP":
            A0 = randn(B, num_ineq, n) * (rand(B, num_ineq, n) < sp_)
            zl, zu = -rand(B, num_ineq, 1), rand(B, num_ineq, 1)
            mi, me = num_ineq, 0
        else:
            A0 = randn(B, num_eq, n) * (rand(B, num_eq, n) < sp_)
            zl = randn(B, num_eq, 1)
            zu = zl.clone()
            mi, me = 0, num_eq
        return dict(Q=(2 * Q0).contiguous(), p=p, A0=A0.contiguous(), zl=zl, zu=zu, num_var=n, num_ineq=mi, num_eq=me)
    if family == "SVM":
        nv, mi = num_var, num_ineq
        n = nv + mi
        sp_ = density if density is not None else 0.5
        Q0 = torch.zeros((B, n, n), device=dev)
        Q0[:, :nv, :nv] = torch.eye(nv, device=dev)
        lamb = 1.0 + randn(B, 1, 1)                                  # np.random.normal(1): mean 1, std 1
        p = torch.cat((torch.zeros((B, nv, 1), device=dev), lamb * torch.ones((B, mi, 1), device=dev)), 1)
        half = mi // 2
        b_hat = torch.cat((torch.ones(half, device=dev), -torch.ones(mi - half, device=dev)))
        A_hat = torch.cat((1 / nv + randn(B, half, nv) / nv, -1 / nv + randn(B, mi - half, nv) / nv), 1)
        A_hat = A_hat * (rand(B, mi, nv) < sp_)
        G = torch.cat((b_hat.view(1, mi, 1) * A_hat, -torch.eye(mi, device=dev).expand(B, mi, mi)), 2)
        A0 = torch.cat((G, torch.eye(n, device=dev).expand(B, n, n)), 1).contiguous()
        inf = float("inf")
        zl = torch.cat((torch.full((B, mi, 1), -inf, device=dev), torch.full((B, nv, 1), -inf, device=dev),
                        torch.zeros((B, mi, 1), device=dev)), 1)
        zu = torch.cat((-torch.ones((B, mi, 1), device=dev), torch.full((B, n, 1), inf, device=dev)), 1)
        # main.py treats every row of A0 as an inequality row for this family (num_eq = 0)
        return dict(Q=(2 * Q0).contiguous(), p=p, A0=A0, zl=zl, zu=zu, num_var=n, num_ineq=mi + n, num_eq=0)
    raise ValueError(f"unknown family {family!r}")
