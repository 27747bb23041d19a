"""Writes tests/golden/datasets/: tiny datasets produced by the reference's OWN writer, so the reader of
`iadmm_b200/data.py` and the (num_ineq, num_eq) handling of the drop-in modules are tested on files in exactly the
format a user has on disk -- not on files the repository wrote itself.

Runs the UNMODIFIED /root/reference/generate_data.py (module-level script code, `runpy`) once per problem family in a
scratch directory, with the two imports this container lacks replaced by stubs:
  * `configargparse` -> argparse (the script only uses ArgumentParser / add_argument / parse_known_args),
  * `osqp`           -> a solver object that accepts setup(...) and reports status 'solved' with all-zero x, y.  The OSQP
                        "solved" filter therefore keeps every instance and the stored labels `x`, `y` (QP / QP_RHS only) are
                        zeros; nothing on the path reads them (main.py loads x_gt / y_gt but never uses them).
Everything else -- the distributions, the dict keys, dense numpy vs scipy csc per family, gzip + pickle, the file and
directory names -- is the reference's code (generate_data.py:31-228).  The SVM branch writes to a hard-coded
'E:/gaoxi/OSQP/OSQP-LSTM/datasets/...' path (generate_data.py:219); on Linux that is a relative path, created below
the scratch directory and copied from there.

    python tests/golden/make_dataset_fixtures.py          # needs /root/reference; the fixtures are committed
"""
import argparse
import os
import runpy
import shutil
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "datasets")
REF = "/root/reference/generate_data.py"

FAMILIES = [   # (prob_type, argv, directory the script writes below ./datasets or the SVM path)
    ("QP", ["--num_var", "12", "--num_ineq", "5", "--num_eq", "4", "--data_size", "3"], "datasets/QP_12_5_4"),
    ("QP_RHS", ["--num_var", "12", "--num_ineq", "5", "--num_eq", "4", "--data_size", "3"], "datasets/QP_RHS_12_5_4"),
    ("Random_QP", ["--num_var", "10", "--num_ineq", "6", "--data_size", "3"], "datasets/Random_QP_10_6"),
    ("Equality_QP", ["--num_var", "10", "--num_eq", "4", "--data_size", "3"], "datasets/Equality_QP_10_4"),
    ("SVM", ["--num_var", "6", "--num_ineq", "4", "--data_size", "3"], "E:/gaoxi/OSQP/OSQP-LSTM/datasets/SVM_10_4"),
]


def stub_modules():
    cap = types.ModuleType("configargparse")

    class ArgumentParser(argparse.ArgumentParser):
        def add_argument(self, *a, **kw):
            kw.pop("is_config_file", None)
            return super().add_argument(*a, **kw)
    cap.ArgumentParser = ArgumentParser
    osqp = types.ModuleType("osqp")

    class OSQP:
        def setup(self, P=None, q=None, A=None, l=None, u=None, **kw):       # noqa: E741
            self.n, self.m = P.shape[0], A.shape[0]

        def solve(self):
            return types.SimpleNamespace(info=types.SimpleNamespace(status="solved"), x=np.zeros(self.n), y=np.zeros(self.m))
    osqp.OSQP = OSQP
    sys.modules["configargparse"], sys.modules["osqp"] = cap, osqp


def main():
    stub_modules()
    shutil.rmtree(OUT, ignore_errors=True)
    argv0, cwd0 = sys.argv, os.getcwd()
    for seed, (family, argv, rel) in enumerate(FAMILIES):
        with tempfile.TemporaryDirectory() as tmp:
            os.chdir(tmp)
            os.makedirs(rel, exist_ok=True)
            np.random.seed(100 + seed)
            torch.manual_seed(100 + seed)
            sys.argv = ["generate_data.py", "--prob_type", family] + argv
            try:
                runpy.run_path(REF, run_name="__main__")
            finally:
                sys.argv = argv0
                os.chdir(cwd0)
            dst = os.path.join(OUT, os.path.basename(rel))
            shutil.copytree(os.path.join(tmp, rel), dst)
            print(family, sorted(os.listdir(dst)), sum(os.path.getsize(os.path.join(dst, f)) for f in os.listdir(dst)), "bytes")


if __name__ == "__main__":
    main()
