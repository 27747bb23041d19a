"""On-chip-resident vs streaming variant over the batch size at the config-1 shape (n=100, 50+50, h=64, K=100)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "i-admm-lstm_b200"))
import torch
import iadmm_b200 as ia
from bench import device_qp_batch
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
n, mi, me, h, K = 100, 50, 50, 64, 100
torch.manual_seed(3)
model = ia.LSTM(None, 2, h, K, dev)
def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
rows = []
for B in (1, 8, 64, 148, 296, 1024, 4096, 16384):
    Q, p, A0, zl, zu = device_qp_batch(B, n, mi, me, 5, dev)
    with torch.no_grad():
        r = timeit(lambda: model.solve(K, mi, me, Q, p, A0, zl, zu, 6e-6))
        s = timeit(lambda: model.solve(K, mi, me, Q, p, A0, zl, zu, 6e-6, streaming=True))
    rows.append(dict(batch=B, resident_ms=r, streaming_ms=s, resident_solves_per_s=B / r * 1e3, streaming_solves_per_s=B / s * 1e3))
    print(json.dumps(rows[-1]), flush=True)
print(json.dumps({"workload": "n=100, 50+50, h=64, K=100, traces on, tc_f16f8", "rows": rows}))
