"""Shared helpers for the test-suite (fixture loading, error metrics)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GATES = ("i", "f", "o", "u")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def t(a, dtype=torch.float32, device="cpu"):
    return torch.as_tensor(np.asarray(a)).to(dtype).to(device).contiguous()


def golden_params(blob, dtype=torch.float32, device="cpu"):
    return {k[4:]: t(v, dtype, device) for k, v in blob.items() if k.startswith("prm_")}


def golden_qp(blob, dtype=torch.float32, device="cpu"):
    return {k[3:]: t(v, dtype, device) for k, v in blob.items() if k.startswith("qp_")}


def rel_err(a, b):
    """||a-b||_2 / ||b||_2 over the whole tensor (per-quantity relative error of north_star)."""
    a = torch.as_tensor(np.asarray(a)).double().cpu() if not torch.is_tensor(a) else a.double().cpu()
    b = torch.as_tensor(np.asarray(b)).double().cpu() if not torch.is_tensor(b) else b.double().cpu()
    fin = torch.isfinite(b)
    assert torch.equal(torch.isfinite(a), fin), "non-finite pattern differs"
    if not bool(fin.all()):
        assert torch.equal(a[~fin], b[~fin])
    a, b = a[fin], b[fin]
    den = float(torch.linalg.vector_norm(b))
    num = float(torch.linalg.vector_norm(a - b))
    return num / den if den > 0 else num
