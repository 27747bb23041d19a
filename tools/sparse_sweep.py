"""KKT-phase time of the bitmap-slab sparse form against the dense form over a density sweep (n = m = 1000): where the
crossover lies decides iadmm_b200.lstm.SPARSE_AUTO_DENSITY.

    python tools/sparse_sweep.py [--batch 128] > gpurun_out/sparse_sweep.json
"""
import argparse, json, os, sys
from ctypes import byref, c_double, c_int
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "i-admm-lstm_b200"))
import torch
import iadmm_b200 as ia

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128); ap.add_argument("--n", type=int, default=1000)
ap.add_argument("--hidden", type=int, default=64); ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
dev = torch.device("cuda:0")
B, n, m, h, K = a.batch, a.n, a.n, a.hidden, a.iters
L = ia.lib()
model = ia.LSTM(None, 2, h, K, dev)
rows = []
for density in (0.001, 0.01, 0.05, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.75):
    g = torch.Generator(device=dev).manual_seed(1)
    Q = torch.randn((B, n, n), device=dev, generator=g) * (torch.rand((B, n, n), device=dev, generator=g) < density)
    Q = (Q + Q.mT) * 0.5 + torch.eye(n, device=dev)
    A0 = torch.randn((B, m, n), device=dev, generator=g) * (torch.rand((B, m, n), device=dev, generator=g) < density)
    p = torch.randn((B, n, 1), device=dev, generator=g)
    zl, zu = -torch.rand((B, m, 1), device=dev, generator=g), torch.rand((B, m, 1), device=dev, generator=g)
    q_sp, a_sp = ia.SparseBatch.pack(Q), ia.SparseBatch.pack(A0)
    out = {"density_A0": a_sp.density, "density_Q": q_sp.density,
           "bytes_per_pass_dense": 4 * (n * n + m * n), "bytes_per_pass_sparse": q_sp.bytes_per_instance + a_sp.bytes_per_instance}
    res = {}
    for tag, kw in (("dense", dict(streaming=True)), ("sparse", dict(sparse=(q_sp, a_sp)))):
        with torch.no_grad():
            for _ in range(2):
                r = model.solve(K, m, 0, Q, p, A0, zl, zu, 6e-6, **kw)
            torch.cuda.synchronize()
            ia._lib.check(L.iadmm_profile_begin(3 * K))
            for _ in range(3):
                r = model.solve(K, m, 0, Q, p, A0, zl, zu, 6e-6, **kw)
            kk, gg, tt, it = c_double(), c_double(), c_double(), c_int()
            ia._lib.check(L.iadmm_profile_end(byref(kk), byref(gg), byref(tt), byref(it)))
        out[tag + "_kkt_ms_per_iteration"] = kk.value / max(1, it.value)
        res[tag] = r
    out["identical"] = bool(torch.equal(res["dense"].x, res["sparse"].x) and torch.equal(res["dense"].pri, res["sparse"].pri))
    out["speedup"] = out["dense_kkt_ms_per_iteration"] / out["sparse_kkt_ms_per_iteration"]
    out["sparse_GBps"] = 2 * B * out["bytes_per_pass_sparse"] / out["sparse_kkt_ms_per_iteration"] / 1e6
    rows.append(out)
    print(json.dumps(out), flush=True)
