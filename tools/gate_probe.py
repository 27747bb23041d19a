"""Development aid: device time of the gate kernel (and the other phases) at config-2 size for the current
environment switches (IADMM_TC_EPI_WARPS, IADMM_TC_EXP, ...), plus a short parity check against the fp32
CUDA-core path.  One process per variant (the switches are read once):

    IADMM_TC_EPI_WARPS=8 python tools/gate_probe.py
"""
import os, sys, json
from ctypes import byref, c_double, c_int
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "i-admm-lstm_b200"))
import torch
import iadmm_b200 as ia
from bench import device_qp_batch, ClockSampler

dev = "cuda:0"
B = int(os.environ.get("PROBE_B", 256)); h = int(os.environ.get("PROBE_H", 800)); K = int(os.environ.get("PROBE_K", 20))
n, mi, me = 1000, 500, 500
mode = os.environ.get("PROBE_MODE", "tc_f16f8")
tag = {k: v for k, v in os.environ.items() if k.startswith("IADMM_")}
L = ia.lib()
torch.manual_seed(17)
out = {"env": tag, "mode": mode, "B": B, "h": h, "K": K}
with torch.no_grad():
    if not os.environ.get("IADMM_TC_EXP"):
        # parity: K=3 on 2 instances against the fp32 CUDA-core path (same weights)
        Qs, ps, As, zls, zus = device_qp_batch(2, n, mi, me, 5, dev)
        res = {}
        for md in ("simt_fp32", mode):
            torch.manual_seed(17)
            mdl = ia.LSTM(None, 2, h, 4, dev, gate_mode=md).eval()
            res[md] = mdl.solve(3, mi, me, Qs, ps, As, zls, zus, 6e-6)
        torch.cuda.synchronize()
        errs = {}
        for k in ("x", "y", "z", "H", "C", "pri", "dual"):
            a, b = getattr(res[mode], k).double(), getattr(res["simt_fp32"], k).double()
            errs[k] = float((a - b).norm() / b.norm())
        out["parity_vs_simt_K3"] = errs
    model = ia.LSTM(None, 2, h, K, dev, gate_mode=mode).eval()
    Q, p, A0, zl, zu = device_qp_batch(B, n, mi, me, 17, dev)
    for _ in range(2):
        r = model.solve(K, mi, me, Q, p, A0, zl, zu, 6e-6)
    torch.cuda.synchronize()
    reps = int(os.environ.get("PROBE_REPS", 15))
    sampler = ClockSampler(torch.device(dev)); sampler.start()
    ia._lib.check(L.iadmm_profile_begin(reps * K))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = model.solve(K, mi, me, Q, p, A0, zl, zu, 6e-6)
    e1.record(); torch.cuda.synchronize()
    kkt, gate, tail, nit = c_double(), c_double(), c_double(), c_int()
    ia._lib.check(L.iadmm_profile_end(byref(kkt), byref(gate), byref(tail), byref(nit)))
    out["clocks"] = sampler.stop()
    out.update(gate_ms=gate.value / nit.value, kkt_ms=kkt.value / nit.value, tail_ms=tail.value / nit.value,
               solve_ms_per_iter=e0.elapsed_time(e1) / (reps * K), finite=bool(torch.isfinite(r.x).all()))
print(json.dumps(out), flush=True)
