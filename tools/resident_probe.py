"""ncu target: config-1 solve on the on-chip-resident kernel (one launch inside the profiler range)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "i-admm-lstm_b200"))
import torch
import iadmm_b200 as ia
from oracle import iadmm_oracle as orc
dev = "cuda:0"
B = int(os.environ.get("B", "64")); n, mi, me, h, K = 100, 50, 50, 64, 100
qp = {k: v.to(dev) for k, v in orc.qp_instances(B, n, mi, me, seed=17).items()}
prm = orc.lstm_parameters(h, K, seed=17)
model = ia.LSTM(None, 2, h, K, dev, gate_mode=os.environ.get("MODE", "tc_f16f8"))
with torch.no_grad():
    for k, v in prm.items(): getattr(model, k).copy_(v.to(dev))
    args = (K, mi, me, qp["Q"], qp["p"], qp["A0"], qp["zl"], qp["zu"], 6e-6)
    for _ in range(3): model.solve(*args)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    model.solve(*args)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("ok")
