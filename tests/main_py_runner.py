"""Runs the reference's UNMODIFIED driver script `main.py` in test mode (`--test --save_sol`, main.py:549-1268) on the small
datasets of tests/golden/datasets/, either on the reference's own modules (build container, CPU: the golden run of
tests/golden/make_main_py_golden.py) or on the drop-in modules of `iadmm_b200` (GPU box: tests/test_gpu_main_py.py).
TEST INFRASTRUCTURE ONLY.

main.py is module-level script code, so it is executed with `runpy` inside a scratch working directory that holds
`./datasets/<family dir>/` and the checkpoint main.py loads (`./results/lstm/params/<name>.pth`, a `state_dict` with the 16
keys of models/lstm.py:21-41).  `configargparse` is absent from this image and is replaced by a thin argparse stub (main.py
uses ArgumentParser / add_argument / parse_known_args only; no --config file is passed).  The modules main.py imports
(`methods.scaling`, `models.lstm`, `models.lu`, `utils`) are placed in `sys.modules` for the duration of the run:
  arm "reference": the reference's files themselves (imported from `ref_root`),
  arm "dropin":    shims that re-export `iadmm_b200.{LSTM, LU, Scaling, primal_dual_loss, obj_fn, ineq_dist, eq_dist, lb_dist,
                   ub_dist}`; `EarlyStopping` and `aug_lagr` (host glue outside the path) stay the reference's.
Returns the dict main.py writes with scipy.io.savemat (`x`, `objs`, `ls_res`, `primal_res`, `dual_res`, `*_fr` with
--feas_rest, ...).

`train=dict(epochs=, lr=, TL=)` runs the TRAINING branch instead (main.py:187-547: TBPTT windows under autograd, Adam,
validation loop, EarlyStopping checkpoint) and returns the checkpoint main.py saved.  Two things have to be supplied from
outside for that, in the shim layer, not in main.py:
  * main.py:191 reads `args.weight_decay`, which main.py never declares (configs/QP.yaml has the key, but configargparse hands
    undeclared config keys to argparse as unknown arguments, which parse_known_args drops): as published the script cannot
    reach its training loop.  The parser stub declares `--weight_decay` with the YAML's value 0.0.
  * main.py seeds `random` only, so the weights `LSTM(...)` draws depend on the device's generator.  Both arms construct the
    model through a subclass that copies the given `prm` in after the stock constructor, so the two runs start from the same
    weights.
"""
import argparse
import contextlib
import importlib.util
import io
import os
import runpy
import shutil
import sys
import types

import scipy.io as sio
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
DATASETS = os.path.join(HERE, "golden", "datasets")
SHADOWED = ("configargparse", "methods", "methods.scaling", "models", "models.lstm", "models.lu", "utils")

# family -> (dataset dir, main.py size flags, checkpoint name, results name) for hidden_dim h and K iterations; the names are
# the ones main.py:552-575 / :1166-1190 build (note: num_eq BEFORE num_ineq for QP / QP_RHS)
CASES = {
    "QP": dict(dir="QP_12_5_4", sizes=["--num_var", "12", "--num_ineq", "5", "--num_eq", "4"],
               ckpt="QP_12_4_5_{K}_{h}.pth", results="QP_12_4_5_{K}_{h}_results.mat"),
    "QP_RHS": dict(dir="QP_RHS_12_5_4", sizes=["--num_var", "12", "--num_ineq", "5", "--num_eq", "4"],
                   ckpt="QP_RHS_12_4_5_{K}_{h}.pth", results="QP_RHS_12_4_5_{K}_{h}_results.mat"),
    "Random_QP": dict(dir="Random_QP_10_6", sizes=["--num_var", "10", "--num_ineq", "6", "--num_eq", "0"],
                      ckpt="Random_QP_10_6_{K}_{h}.pth", results="Random_QP_10_6_{K}_{h}_results.mat"),
    "Equality_QP": dict(dir="Equality_QP_10_4", sizes=["--num_var", "10", "--num_ineq", "0", "--num_eq", "4"],
                        ckpt="Equality_QP_10_4_{K}_{h}.pth", results="Equality_QP_10_4_{K}_{h}_results.mat"),
    "SVM": dict(dir="SVM_10_4", sizes=["--num_var", "10", "--num_ineq", "4", "--num_eq", "0"],
                ckpt="SVM_10_4_{K}_{h}.pth", results="SVM_10_4_{K}_{h}.pth"),      # (main.py:1184 names the SVM results .pth)
}


def _configargparse_stub():
    mod = types.ModuleType("configargparse")

    class ArgumentParser(argparse.ArgumentParser):
        def add_argument(self, *a, **kw):
            kw.pop("is_config_file", None)
            return super().add_argument(*a, **kw)

        def parse_known_args(self, *a, **kw):
            if not any(act.dest == "weight_decay" for act in self._actions):      # see the module docstring
                super().add_argument("--weight_decay", type=float, default=0.0)
            return super().parse_known_args(*a, **kw)
    mod.ArgumentParser = ArgumentParser
    return mod


def _load_file(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _with_weights(cls, prm):
    """`cls` with the given parameter tensors copied in after the stock constructor (training runs: same start on any device)."""
    if prm is None:
        return cls

    class LSTM(cls):
        def __init__(self, *a, **kw):
            super().__init__(*a, **kw)
            with torch.no_grad():
                for k, v in prm.items():
                    getattr(self, k).copy_(v.to(getattr(self, k).device))
    return LSTM


def _arm_modules(arm, ref_root, init_prm=None):
    mods = {"configargparse": _configargparse_stub()}
    ref_utils = _load_file("_main_py_ref_utils", os.path.join(ref_root, "utils.py"))
    if arm == "reference":
        lstm = _load_file("_main_py_ref_lstm", os.path.join(ref_root, "models", "lstm.py"))
        lu = _load_file("_main_py_ref_lu", os.path.join(ref_root, "models", "lu.py"))
        scaling = _load_file("_main_py_ref_scaling", os.path.join(ref_root, "methods", "scaling.py"))
        utils = ref_utils
        if init_prm is not None:      # (a separate module object: models/lstm.py:13 looks its own class up by its global name)
            stock, lstm = lstm.LSTM, types.ModuleType("models.lstm")
            lstm.LSTM = _with_weights(stock, init_prm)
    else:
        import iadmm_b200 as ia
        lstm, lu, scaling, utils = (types.ModuleType(n) for n in ("models.lstm", "models.lu", "methods.scaling", "utils"))
        lstm.LSTM, lu.LU, scaling.Scaling = _with_weights(ia.LSTM, init_prm), ia.LU, ia.Scaling
        for name in ("primal_dual_loss", "obj_fn", "ineq_dist", "eq_dist", "lb_dist", "ub_dist"):
            setattr(utils, name, getattr(ia, name))
        utils.EarlyStopping, utils.aug_lagr = ref_utils.EarlyStopping, ref_utils.aug_lagr
    models, methods = types.ModuleType("models"), types.ModuleType("methods")
    models.lstm, models.lu, methods.scaling = lstm, lu, scaling
    mods.update({"models": models, "models.lstm": lstm, "models.lu": lu, "methods": methods, "methods.scaling": scaling,
                 "utils": utils})
    return mods


def run_main_py(family, arm, ref_root, workdir, prm, h, K, device, scaling=True, feas_rest=0, batch=3, train=None, case=None,
                data_size=3):
    """One `python main.py ... --test --save_sol` run; returns (results dict, captured stdout).  With `train`: one training
    run, returns (saved checkpoint, captured stdout).  `case` overrides the CASES entry (a dataset the caller has already put
    below `workdir/datasets/`, tools/main_py_speed.py)."""
    case = case or CASES[family]
    os.makedirs(os.path.join(workdir, "datasets"), exist_ok=True)
    dst = os.path.join(workdir, "datasets", case["dir"])
    if not os.path.isdir(dst):
        shutil.copytree(os.path.join(DATASETS, case["dir"]), dst)
    params_dir = os.path.join(workdir, "results", "lstm", "params")
    os.makedirs(params_dir, exist_ok=True)
    if not train:
        torch.save({k: v.clone() for k, v in prm.items()}, os.path.join(params_dir, case["ckpt"].format(K=K, h=h)))
    argv = ["main.py", "--model_name", "LSTM", "--prob_type", family, *case["sizes"], "--input_dim", "2", "--hidden_dim", str(h),
            "--outer_T", str(K), "--test_outer_T", str(K), "--truncated_length", str(K), "--sigma", "0.000006", "--data_size", str(data_size),
            "--val_frac", "0", "--test_frac", "1", "--batch_size", str(batch), "--test_batch_size", str(batch), "--device", device,
            "--save_dir", "./results/", "--seed", "17", "--test", "--save_sol", "--eq_tol", "0.2", "--ineq_tol", "0.2"]
    if train:
        # 3 instances: 1 for training, 1 for validation (train_size = int(3 * 0.66), val_size = int(3 * 0.34)); the tolerances are
        # wide open so EarlyStopping.step (utils.py:15-44) saves whenever the validation objective improves
        argv = [a for a in argv if a not in ("--test", "--save_sol")]
        argv += ["--num_epoch", str(train["epochs"]), "--lr", str(train["lr"]), "--early_stop_mode", "min", "--patience", "100"]
        for flag, val in (("--val_frac", "0.34"), ("--test_frac", "0"), ("--batch_size", "1"), ("--truncated_length", str(train["TL"])),
                          ("--eq_tol", "1e9"), ("--ineq_tol", "1e9")):
            argv[argv.index(flag) + 1] = val
    if scaling:
        argv.append("--scaling")
    if feas_rest:
        argv += ["--feas_rest", "--feas_rest_num", str(feas_rest)]
    saved = {k: sys.modules.get(k) for k in SHADOWED}
    argv0, cwd0 = sys.argv, os.getcwd()
    out = io.StringIO()
    try:
        sys.modules.update(_arm_modules(arm, ref_root, prm if train else None))
        sys.argv = argv
        os.chdir(workdir)
        with contextlib.redirect_stdout(out):
            runpy.run_path(os.path.join(ref_root, "main.py"), run_name="__main__")
    finally:
        os.chdir(cwd0)
        sys.argv = argv0
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    if train:
        (name,) = os.listdir(params_dir)             # main.py:78-96 names it (num_ineq before num_eq here)
        return torch.load(os.path.join(params_dir, name), map_location="cpu"), out.getvalue()
    res = sio.loadmat(os.path.join(workdir, "results", "lstm", case["results"].format(K=K, h=h)), appendmat=False)
    return {k: v for k, v in res.items() if not k.startswith("__")}, out.getvalue()
