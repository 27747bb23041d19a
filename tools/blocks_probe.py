"""QP family (diagonal Q) at the headline shape: dense vs block-skip form of Q, a few iterations (ncu target).
python tools/blocks_probe.py [dense|blocks] [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "i-admm-lstm_b200"))
import torch
import iadmm_b200 as ia
from bench import device_qp_batch
form = sys.argv[1] if len(sys.argv) > 1 else "blocks"; B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda:0"); n, mi, me, h, K = 1000, 500, 500, 64, 4
Q, p, A0, zl, zu = device_qp_batch(B, n, mi, me, 17, dev)
model = ia.LSTM(None, 2, h, K, dev)
sp = (ia.SparseBatch.blocks(Q), None) if form == "blocks" else None
with torch.no_grad():
    for _ in range(2):
        r = model.solve(K, mi, me, Q, p, A0, zl, zu, 6e-6, sparse=sp, streaming=True)
torch.cuda.synchronize()
print("ok", form, float(r.x.abs().sum()), None if sp is None else sp[0].occupancy)
