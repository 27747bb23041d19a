"""What a user of the reference sees: the reference's UNMODIFIED `main.py --test` (its loader, Ruiz scaling, the literal
per-iteration loop with ~10 host syncs per iteration, un-scaling bmm's, metrics) on a QP dataset of the headline size, once on
the reference's own modules on this GPU (stock PyTorch) and once on the drop-in modules -- nothing but the imports differs
(tests/main_py_runner.py).  Reports main.py's own "Parallel Time" line (seconds per instance: the model() calls + scaling, its
time.time() brackets, main.py:880-890) and the wall time of the whole script.

    python tools/main_py_speed.py [instances batch hidden K n]
"""
import json
import os
import re
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p_ in (ROOT, os.path.join(ROOT, "i-admm-lstm_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p_)
import numpy as np
import torch
from main_py_runner import run_main_py
from iadmm_b200 import data
from oracle.iadmm_oracle import lstm_parameters          # weight generator only


def main():
    inst = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    h = int(sys.argv[3]) if len(sys.argv) > 3 else 800
    K = int(sys.argv[4]) if len(sys.argv) > 4 else 100
    n = int(sys.argv[5]) if len(sys.argv) > 5 else 1000
    mi = me = n // 2
    ref = os.path.join(ROOT, "baseline", "_ref")
    prm = lstm_parameters(h, K, seed=17)
    case = dict(dir=f"QP_{n}_{mi}_{me}", sizes=["--num_var", str(n), "--num_ineq", str(mi), "--num_eq", str(me)],
                ckpt=f"QP_{n}_{me}_{mi}_{{K}}_{{h}}.pth", results=f"QP_{n}_{me}_{mi}_{{K}}_{{h}}_results.mat")
    out = {"instances": inst, "test_batch_size": batch, "hidden_dim": h, "K": K, "n": n}
    with tempfile.TemporaryDirectory() as tmp:
        d = os.path.join(tmp, "datasets", case["dir"])
        qp = data.generate_qp_batch(inst, n, mi, me, seed=5, device="cuda:0", as_stored=True)
        for i in range(inst):                                   # the reference's on-disk format (generate_data.py:77-92)
            rec = {k: qp[k][i].cpu().numpy() for k in ("Q", "p", "G", "c", "A", "b", "A0", "zl", "zu")}
            rec.update(x=np.zeros(n), y=np.zeros(mi + me))
            data.write_instance(os.path.join(d, f"qp_{i}.gz"), rec)
        res = {}
        for arm in ("dropin", "reference", "dropin"):          # (the first drop-in run warms the allocator and the library up)
            torch.cuda.synchronize()
            t0 = time.time()
            r, log = run_main_py("QP", arm, ref, tmp, prm, h, K, "cuda:0", scaling=True, batch=batch, case=case, data_size=inst)
            torch.cuda.synchronize()
            wall = time.time() - t0
            par = float(re.findall(r"Parallel Time : ([0-9.eE+-]+)", log)[-1])
            res[arm] = r
            out[arm] = {"main_py_parallel_time_s_per_instance": par, "solves_per_s_by_main_py_clock": 1.0 / par,
                        "script_wall_s": round(wall, 2)}
        a, b = np.asarray(res["dropin"]["x"], dtype=np.float64), np.asarray(res["reference"]["x"], dtype=np.float64)
        out["x_rel_err_dropin_vs_reference_on_this_gpu"] = float(np.linalg.norm(a - b) / np.linalg.norm(b))
        out["speedup_by_main_py_clock"] = out["reference"]["main_py_parallel_time_s_per_instance"] / out["dropin"]["main_py_parallel_time_s_per_instance"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
