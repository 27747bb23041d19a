"""Ruiz equilibration: device time and HBM traffic class of the chain form (read-only norm passes on the original matrices,
one write) against the in-place form of round 1 (development switch IADMM_RUIZ_CHAIN=0, development build), and bit-identity
of the two.  Run once per form:  IADMM_B200_LIB=.../libiadmm_b200_dev.so [IADMM_RUIZ_CHAIN=0] python tools/ruiz_ab.py"""
import hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "i-admm-lstm_b200"))
import torch
import iadmm_b200 as ia
from bench import device_qp_batch
dev = torch.device("cuda:0")
B, n, mi, me = int(os.environ.get("RZ_B", 256)), int(os.environ.get("RZ_N", 1000)), 0, 0
mi = me = n // 2
Q, p, A0, zl, zu = device_qp_batch(B, n, mi, me, 17, dev)
sc = ia.Scaling(n, mi + me, 10, dev)
for _ in range(3):
    out = sc.scale_data(Q, p, A0, zl, zu)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    out = sc.scale_data(Q, p, A0, zl, zu)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
h = hashlib.sha256()
for t in list(out) + [sc.d, sc.e, sc.c_vec]:
    h.update(t.cpu().numpy().tobytes())
mat = 4.0 * B * (n * n + (mi + me) * n)
print(json.dumps({"form": ("in-place (round 1)" if os.environ.get("IADMM_RUIZ_CHAIN") == "0" else "chain")
                  + (", vector kernel walks the chunk partials" if os.environ.get("IADMM_RUIZ_FOLD") == "0" else ""), "B": B, "n": n, "ms": ms,
                  "passes_equivalent_at_6559GBps": ms * 1e-3 * 6559.4e9 / mat, "algorithmic_passes": 11, "sha256": h.hexdigest()}))
