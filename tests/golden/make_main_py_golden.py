"""Golden results of the reference's UNMODIFIED driver: `python main.py ... --test --save_sol` (main.py:549-1268) on the
reference's own modules (CPU), for the five dataset fixtures of tests/golden/datasets/ -- QP additionally with Stage II
(`--feas_rest`, models/lu.py) -- and of one run of its TRAINING branch (main.py:187-547, QP fixture).  tests/test_gpu_main_py.py runs the same unmodified script on the drop-in modules and compares
what main.py saves with scipy.io.savemat and what it prints.

    python tests/golden/make_main_py_golden.py             # build container only (/root/reference)
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from main_py_runner import run_main_py            # noqa: E402
from oracle.iadmm_oracle import lstm_parameters   # noqa: E402  (weight generator only)

H, K = 8, 5
RUNS = [("QP", 4), ("QP_RHS", 0), ("Random_QP", 0), ("Equality_QP", 0), ("SVM", 0)]      # (family, --feas_rest_num)


def main():
    for i, (family, fr) in enumerate(RUNS):
        prm = lstm_parameters(H, K, seed=50 + i, scale=8.0)
        with tempfile.TemporaryDirectory() as tmp:
            res, out = run_main_py(family, "reference", "/root/reference", tmp, prm, H, K, "cpu", scaling=True, feas_rest=fr)
        keep = {k: np.asarray(v, dtype=np.float64) for k, v in res.items() if k != "time" and np.asarray(v).size}
        np.savez_compressed(os.path.join(HERE, f"main_py_{family}.npz"), meta=np.array([H, K, fr, 50 + i]), stdout=np.array(out),
                            **{"res_" + k: v for k, v in keep.items()}, **{"prm_" + k: v.numpy() for k, v in prm.items()})
        print(family, sorted(keep), "| last line:", [l for l in out.splitlines() if l.startswith("Test_")][-1])


def train_case():
    """The TRAINING branch (main.py:187-547) for the QP fixture: 2 epochs of 2 TBPTT windows (outer_T 6, truncated_length 3),
    Adam lr 1e-3, validation + EarlyStopping checkpoint.  See main_py_runner.py for the two things the shim layer supplies."""
    h, K, seed = 8, 6, 60
    prm = lstm_parameters(h, K, seed=seed)
    with tempfile.TemporaryDirectory() as tmp:
        ck, out = run_main_py("QP", "reference", "/root/reference", tmp, prm, h, K, "cpu", scaling=True, train=dict(epochs=2, lr=1e-3, TL=3))
    np.savez_compressed(os.path.join(HERE, "main_py_train_QP.npz"), meta=np.array([h, K, 2, 3, seed]), lr=np.float64(1e-3), stdout=np.array(out),
                        **{"ckpt_" + k: v.numpy() for k, v in ck.items()}, **{"prm_" + k: v.numpy() for k, v in prm.items()})
    print("train QP:", [l for l in out.splitlines() if "Train_Obj" in l])


if __name__ == "__main__":
    train_case()
    main()
