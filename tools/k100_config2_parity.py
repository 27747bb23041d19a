"""north_star's parity statement at the headline shape: K=100 iterations at n=1000, 500+500, hidden_dim=800, --scaling,
GPU path (default F16F8 mode and the fp32 CUDA-core path) against the CPU oracle in fp32 (the reference's arithmetic)
and fp64 (tie-breaker), same inputs and weights.  Prints rel. errors of x^K, y^K, z^K and the residual traces."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "i-admm-lstm_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import iadmm_b200 as ia
from oracle import iadmm_oracle as orc
from helpers import rel_err

B, n, mi, me, h, K = int(os.environ.get("PB", 2)), 1000, 500, 500, 800, 100
seed, wscale = int(os.environ.get("PSEED", 41)), float(os.environ.get("PWSCALE", 1.0))
torch.set_num_threads(os.cpu_count())
res = {}
t0 = time.time()
for dt, tag in ((torch.float32, "oracle_fp32"), (torch.float64, "oracle_fp64")):
    qp = orc.qp_instances(B, n, mi, me, seed=seed, dtype=dt)
    prm = orc.lstm_parameters(h, K, seed=seed, scale=wscale, dtype=dt)
    Qs, ps, As, zls, zus, so = orc.ruiz_equilibrate(qp["Q"], qp["p"], qp["A0"], qp["zl"], qp["zu"], 10)
    res[tag] = orc.solve(prm, K, mi, me, Qs, ps, As, zls, zus, 6e-6, h, form="block")
print("oracle runs: %.1f s" % (time.time() - t0), flush=True)
qp = orc.qp_instances(B, n, mi, me, seed=seed)
prm = orc.lstm_parameters(h, K, seed=seed, scale=wscale)
for mode in ("simt_fp32", "tc_f16f8", "tc_3xfp16", "tc_1xfp16"):
    model = ia.LSTM(None, 2, h, K, "cuda:0", gate_mode=mode)
    with torch.no_grad():
        for k, v in prm.items(): getattr(model, k).copy_(v.cuda())
        sc = ia.Scaling(n, mi + me, 10, "cuda:0")
        Q, p, A0, zl, zu = sc.scale_data(*(qp[k].cuda() for k in ("Q", "p", "A0", "zl", "zu")))
        res[mode] = model.solve(K, mi, me, Q, p, A0, zl, zu, 6e-6)
    torch.cuda.synchronize()
out = {}
for a in ("oracle_fp32", "simt_fp32", "tc_f16f8", "tc_3xfp16", "tc_1xfp16"):
    for b in ("oracle_fp64", "oracle_fp32"):
        if a == b: continue
        out[f"{a} vs {b}"] = {k: float("%.2e" % rel_err(getattr(res[a], k).double().cpu(), getattr(res[b], k).double())) for k in ("x", "y", "z", "pri", "dual")}
print(json.dumps({"B": B, "wscale": wscale, "seed": seed, "errors": out}, indent=1))
