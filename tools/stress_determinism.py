"""Stress: the fused solve must be bit-for-bit repeatable (no race in the mbarrier protocols, the bulk-copied parameter
blocks or the row-interleaved layout conversion).  Full-size runs repeated, hashes of every output compared."""
import os, sys, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "i-admm-lstm_b200"))
import torch
import iadmm_b200 as ia
from bench import device_qp_batch
dev = "cuda:0"
def digest(r):
    h = hashlib.sha256()
    for k in ("x", "y", "z", "xv", "H", "C", "pri", "dual"):
        h.update(getattr(r, k).contiguous().cpu().numpy().tobytes())
    return h.hexdigest()[:16]
ok = True
for (B, n, hdim, K, reps) in ((256, 1000, 800, 30, 4), (256, 1000, 208, 30, 4), (37, 260, 64, 50, 6), (3, 100, 48, 20, 10)):
    torch.manual_seed(17)
    model = ia.LSTM(None, 2, hdim, K, dev).eval()
    Q, p, A0, zl, zu = device_qp_batch(B, n, n // 2, n // 2, 17, dev)
    ds = []
    with torch.no_grad():
        for _ in range(reps):
            ds.append(digest(model.solve(K, n // 2, n // 2, Q, p, A0, zl, zu, 6e-6, streaming=True)))
    same = len(set(ds)) == 1
    ok &= same
    print(f"B={B} n={n} h={hdim} K={K}: {reps} runs, digests {'identical' if same else ds}", flush=True)
print("determinism ok" if ok else "NON-DETERMINISTIC")
sys.exit(0 if ok else 1)
