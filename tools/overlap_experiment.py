"""Experiment: solve two half batches on two streams so the HBM-bound KKT passes of one half overlap the
tensor-bound gate kernel of the other (persistent gate grid restricted with IADMM_TC_MAX_SMS)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "i-admm-lstm_b200"))
import torch
import iadmm_b200 as ia
from bench import device_qp_batch
dev = "cuda:0"
n = int(os.environ.get("OV_N", 1000)); h = int(os.environ.get("OV_H", 800)); B = int(os.environ.get("OV_B", 256)); K = int(os.environ.get("OV_K", 100))
mi = me = n // 2
torch.manual_seed(17)
models = [ia.LSTM(None, 2, h, K, dev) for _ in range(2)]
with torch.no_grad():
    for a, b in zip(models[0].parameters(), models[1].parameters()): b.copy_(a)
Q, p, A0, zl, zu = device_qp_batch(B, n, mi, me, 17, dev)
halves = [tuple(t[s].contiguous() for t in (Q, p, A0, zl, zu)) for s in (slice(0, B // 2), slice(B // 2, B))]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
def full():
    return models[0].solve(K, mi, me, Q, p, A0, zl, zu, 6e-6)
def split():
    cur = torch.cuda.current_stream()
    outs = []
    for s, mdl, hv in zip(streams, models, halves):
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            outs.append(mdl.solve(K, mi, me, *hv, 6e-6))
    for s in streams: cur.wait_stream(s)
    return outs
def timeit(fn, reps=3):
    with torch.no_grad():
        fn(); fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): r = fn()
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r
ms_full, rf = timeit(full)
ms_split, rs = timeit(split)
same = torch.equal(rf.x, torch.cat([rs[0].x, rs[1].x])) and torch.equal(rf.pri, torch.cat([rs[0].pri, rs[1].pri], 1))
print(f"n={n} h={h} B={B} K={K} max_sms={os.environ.get('IADMM_TC_MAX_SMS','148')}: full batch {ms_full:.1f} ms ({B/ms_full*1e3:.1f} solves/s)   two half batches on two streams {ms_split:.1f} ms ({B/ms_split*1e3:.1f} solves/s)  identical={same}")
