"""One short sparse solve at a given density (ncu target): python tools/sparse_probe.py DENSITY [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "i-admm-lstm_b200"))
import torch
import iadmm_b200 as ia
density = float(sys.argv[1]); B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device("cuda:0"); n = m = 1000; h, K = 64, 3
g = torch.Generator(device=dev).manual_seed(1)
Q = torch.randn((B, n, n), device=dev, generator=g) * (torch.rand((B, n, n), device=dev, generator=g) < density)
Q = (Q + Q.mT) * 0.5 + torch.eye(n, device=dev)
A0 = torch.randn((B, m, n), device=dev, generator=g) * (torch.rand((B, m, n), device=dev, generator=g) < density)
p = torch.randn((B, n, 1), device=dev, generator=g)
zl, zu = -torch.rand((B, m, 1), device=dev, generator=g), torch.rand((B, m, 1), device=dev, generator=g)
model = ia.LSTM(None, 2, h, K, dev)
sp = (ia.SparseBatch.pack(Q), ia.SparseBatch.pack(A0))
with torch.no_grad():
    for _ in range(2):
        r = model.solve(K, m, 0, Q, p, A0, zl, zu, 6e-6, sparse=sp)
torch.cuda.synchronize()
print("ok", float(r.x.abs().sum()))
