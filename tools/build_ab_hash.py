"""sha256 over the outputs of short solves at several hidden sizes with whatever library IADMM_B200_LIB names: two builds whose
kernels are meant to be bit-identical (e.g. a re-ordered operand layout) must print the same lines."""
import hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "i-admm-lstm_b200"))
import torch
import iadmm_b200 as ia
from bench import device_qp_batch

dev = torch.device("cuda:0")
for h, B, n, K in ((200, 8, 1000, 20), (64, 4, 300, 20), (384, 4, 1000, 10), (400, 4, 1000, 10), (800, 6, 1000, 10), (72, 3, 500, 12)):
    mi = me = n // 2
    torch.manual_seed(3)
    model = ia.LSTM(None, 2, h, K, dev, gate_mode="tc_f16f8")
    Q, p, A0, zl, zu = device_qp_batch(B, n, mi, me, 77, dev)
    sc = ia.Scaling(n, mi + me, 10, dev)
    data = sc.scale_data(Q, p, A0, zl, zu)
    with torch.no_grad():
        r = model.solve(K, mi, me, *data, 6e-6, scaling=sc, streaming=True)
    torch.cuda.synchronize()
    hsh = hashlib.sha256()
    for k in ("x", "y", "z", "xv", "H", "C", "pri", "dual"):
        hsh.update(getattr(r, k).contiguous().cpu().numpy().tobytes())
    print(json.dumps({"hidden": h, "batch": B, "n": n, "K": K, "sha256": hsh.hexdigest()[:32]}), flush=True)
