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
    when present, G, c, A, b, lb, ub.  Also returns the sizes main.py derives (num_var, num_ineq, num_eq)."""
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
