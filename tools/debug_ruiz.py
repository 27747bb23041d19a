import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "i-admm-lstm_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, numpy as np
import iadmm_b200 as ia
from oracle import iadmm_oracle as orc
from helpers import load_golden, golden_qp
def ulps(a, b):
    a = a.cpu().contiguous().view(torch.int32).long(); b = b.cpu().contiguous().view(torch.int32).long()
    d = (a - b).abs(); return int(d.max()), int((d > 0).sum()), d.numel()
for name in ["ruiz_small", "ruiz_c1"]:
    g = load_golden(name); cpu = golden_qp(g); qp = {k: v.cuda() for k, v in cpu.items()}
    B, n, mi, me, _ = (int(v) for v in g["meta"])
    for ites in (1, 2, 10):
        Qo, po, Ao, zlo, zuo, so = orc.ruiz_equilibrate(cpu["Q"], cpu["p"], cpu["A0"], cpu["zl"], cpu["zu"], ites)
        sc = ia.Scaling(n, mi + me, ites, "cuda:0")
        Q, p, A0, zl, zu = sc.scale_data(qp["Q"], qp["p"], qp["A0"], qp["zl"], qp["zu"])
        fin = torch.isfinite(zlo)
        print(name, ites, "d", ulps(sc.d, so.d), "e", ulps(sc.e, so.e), "c", ulps(sc.c_vec, so.c.reshape(-1)),
              "A0", ulps(A0, Ao), "Q", ulps(Q, Qo), "p", ulps(p, po), "zu", ulps(zu, zuo), "zl", ulps(zl.cpu()[fin], zlo[fin]))
        if ites == 1:
            dd = (sc.d.cpu() != so.d).nonzero()[:3]
            for ij in dd:
                b_, j = int(ij[0]), int(ij[1])
                col = torch.maximum(cpu["Q"][b_].abs().amax(0), cpu["A0"][b_].abs().amax(0))[j]
                print("  d diff at", b_, j, float(sc.d[b_, j]), float(so.d[b_, j]), "colnorm", float(col), "1/sqrt", float(1 / torch.sqrt(col)))
