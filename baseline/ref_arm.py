"""Driver of the UNMODIFIED reference modules (models/lstm.py, methods/scaling.py, utils.py) for the reference arms of
bench.py and for the on-device parity tests.  MEASUREMENT / TEST INFRASTRUCTURE ONLY -- nothing under i-admm-lstm_b200/
imports this package.

The reference is a Python package and `/root/reference` does not exist on the GPU box, so `__graft_entry__.build()`
copies the four files of the path verbatim into the git-ignored `baseline/_ref/` (never into history); gpurun ships that
directory with the snapshot.  This module only *drives* them the way main.py does:

    main.py:818-834  Scaling(...).scale_data(Q, p, A0, zl, zu)
    main.py:837-843  zero state
    main.py:874-887  for t in range(K): model(t, ...)
    main.py:346/:955 primal_dual_loss after every iteration

`available()` says whether the copy is there; every entry point raises a clear error otherwise (callers fall back to the
oracle port for the CPU arm and report `unavailable` for the GPU arm).
"""
import importlib
import os
import shutil
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REF_FILES = ("models/lstm.py", "models/lu.py", "methods/scaling.py", "utils.py")
DRIVER_FILES = ("main.py",)      # the reference's driver script, run UNMODIFIED on the drop-in modules by tests/test_gpu_main_py.py
SIGMA = 6e-6          # configs/QP.yaml:14


def install(src="/root/reference"):
    """Copy the reference's files of the path into baseline/_ref (build container only).  Returns True if present after."""
    if os.path.isdir(src):
        for rel in REF_FILES + DRIVER_FILES:
            dst = os.path.join(REF_DIR, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(os.path.join(src, rel), dst)
    return available()


def available():
    return all(os.path.exists(os.path.join(REF_DIR, rel)) for rel in REF_FILES)


_mods = None


def modules():
    """(LSTM, Scaling, primal_dual_loss) of the reference, imported from baseline/_ref under private module names so they
    cannot shadow (or be shadowed by) this repo's own `utils` / `models`."""
    global _mods
    if _mods is None:
        if not available():
            raise RuntimeError("baseline/_ref is missing: run __graft_entry__.build() in the build container "
                               "(it copies the reference's models/, methods/, utils.py there)")
        def load(name, rel):
            spec = importlib.util.spec_from_file_location("_iadmm_ref_" + name, os.path.join(REF_DIR, rel))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod
        import importlib.util  # noqa: F401
        lstm = load("lstm", "models/lstm.py")
        scaling = load("scaling", "methods/scaling.py")
        utils = load("utils", "utils.py")
        _mods = (lstm.LSTM, scaling.Scaling, utils.primal_dual_loss)
    return _mods


def make_model(prm, h, K, device, dtype=torch.float32):
    """The reference LSTM with the given 16 parameter tensors (state_dict names of models/lstm.py:21-41)."""
    LSTM, _, _ = modules()
    model = LSTM(None, 2, h, K, device)
    with torch.no_grad():
        for k, v in prm.items():
            getattr(model, k).copy_(v.to(device))
    if dtype == torch.float64:
        model = model.double()
    return model.to(device).eval()


def solve(model, K, num_ineq, num_eq, Q, p, A0, zl, zu, sigma=SIGMA, scaling_ites=10, traces=True):
    """The reference's test-mode solve on the tensors' device: optional Ruiz scaling, zero state, K calls of model(t, ...),
    primal_dual_loss on the solve's data after each.  Returns a dict of the final iterates and [K,B] traces."""
    _, Scaling, primal_dual_loss = modules()
    dev, dt = Q.device, Q.dtype
    B, n = Q.shape[0], Q.shape[1]
    m = num_ineq + num_eq
    h = model.hidden_dim
    sc = None
    if scaling_ites:
        sc = Scaling(n, m, scaling_ites, dev)
        Q, p, A0, zl, zu = sc.scale_data(Q, p, A0, zl, zu)
    x = torch.zeros((B, n, 1), device=dev, dtype=dt); y = torch.zeros((B, m, 1), device=dev, dtype=dt)
    z = torch.zeros((B, m, 1), device=dev, dtype=dt); xv = torch.zeros((B, n + m, 1), device=dev, dtype=dt)
    H = torch.zeros((B, n + m, h), device=dev, dtype=dt); C = torch.zeros((B, n + m, h), device=dev, dtype=dt)
    pri, dual = [], []
    with torch.no_grad():
        for t in range(K):
            x, y, z, xv, H, C, _, _, _ = model(t, num_ineq, num_eq, x, y, z, xv, sigma, H, C, Q=Q, p=p, A0=A0, lb=None, ub=None,
                                               zl=zl, zu=zu)
            if traces:
                pr, du, _ = primal_dual_loss(x, y, z, Q, p, A0)
                pri.append(pr.reshape(B)); dual.append(du.reshape(B))
    out = dict(x=x, y=y, z=z, xv=xv, H=H, C=C, scaling=sc, data=(Q, p, A0, zl, zu))
    if traces:
        out["pri"] = torch.stack(pri); out["dual"] = torch.stack(dual)
    return out


if __name__ == "__main__":
    print("installed" if install() else "reference not available")
