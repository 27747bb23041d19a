"""Small end-to-end run of every kernel family for compute-sanitizer (memcheck / initcheck one at a time)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "i-admm-lstm_b200"))
import torch
import iadmm_b200 as ia
from oracle import iadmm_oracle as orc
dev = "cuda:0"
for (B, n, mi, me, h, K) in ((3, 37, 9, 11, 16, 3), (2, 132, 40, 29, 48, 3), (1, 20, 0, 0, 16, 2)):
    qp = {k: v.to(dev) for k, v in orc.qp_instances(B, n, mi, me, seed=3).items()} if mi + me > 0 else None
    if qp is None:
        qp = dict(Q=torch.eye(n, device=dev).repeat(B, 1, 1), p=torch.rand((B, n, 1), device=dev), A0=torch.zeros((B, 0, n), device=dev),
                  zl=torch.zeros((B, 0, 1), device=dev), zu=torch.zeros((B, 0, 1), device=dev))
    prm = orc.lstm_parameters(h, K, seed=3)
    sc = ia.Scaling(n, mi + me, 10, dev)
    Q, p, A0, zl, zu = sc.scale_data(qp["Q"], qp["p"], qp["A0"], qp["zl"], qp["zu"])
    for mode in ("simt_fp32", "tc_3xfp16", "tc_f16f8", "tc_1xfp16"):
        model = ia.LSTM(None, 2, h, K, dev, gate_mode=mode)
        with torch.no_grad():
            for k, v in prm.items(): getattr(model, k).copy_(v.to(dev))
            r = model.solve(K, mi, me, Q, p, A0, zl, zu, 6e-6, scaling=sc)
            out = model(0, mi, me, r.x, r.y, r.z, r.xv, 6e-6, r.H, r.C, Q=Q, p=p, A0=A0, lb=None, ub=None, zl=zl, zu=zu)
        torch.cuda.synchronize()
    if mi + me > 0:
        model = ia.LSTM(None, 2, h, K, dev, gate_mode="simt_fp32")
        m = mi + me
        st = [torch.zeros((B, n, 1), device=dev), torch.zeros((B, m, 1), device=dev), torch.zeros((B, m, 1), device=dev),
              torch.zeros((B, n + m, 1), device=dev), torch.zeros((B, n + m, h), device=dev), torch.zeros((B, n + m, h), device=dev)]
        loss = 0.0
        for t in range(2):
            *st, _, _, _ = model(t, mi, me, *st[:4], 6e-6, st[4], st[5], Q=Q, p=p, A0=A0, lb=None, ub=None, zl=zl, zu=zu)
            loss = loss + ia.primal_dual_loss(st[0], st[1], st[2], Q, p, A0)[2].mean()
        loss.backward(); torch.cuda.synchronize()
    print("ok", B, n, mi, me, h, flush=True)
print("sanitize run finished")
