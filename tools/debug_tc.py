"""Development aid: one gate-kernel step in every mode against the fp32 CUDA-core path and the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "i-admm-lstm_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import iadmm_b200 as ia
from oracle import iadmm_oracle as orc
from helpers import rel_err

def run(B, n, mi, me, h, K, scale=1.0, seed=3):
    qp = orc.qp_instances(B, n, mi, me, seed)
    prm = orc.lstm_parameters(h, max(K, 4), seed, scale=scale)
    g = torch.Generator().manual_seed(seed + 1)
    m = mi + me
    st = [torch.randn((B, n, 1), generator=g), torch.randn((B, m, 1), generator=g), torch.randn((B, m, 1), generator=g),
          torch.randn((B, n + m, 1), generator=g), torch.tanh(torch.randn((B, n + m, h), generator=g)), torch.randn((B, n + m, h), generator=g)]
    ref = orc.solve(prm, K, mi, me, qp["Q"], qp["p"], qp["A0"], qp["zl"], qp["zu"], 6e-6, h, state=[s.clone() for s in st], form="block")
    out = {}
    for mode in os.environ.get("MODES", "simt_fp32,tc_3xfp16,tc_f16f8,tc_1xfp16").split(","):
        model = ia.LSTM(None, 2, h, max(K, 4), "cuda:0", gate_mode=mode)
        with torch.no_grad():
            for k, v in prm.items(): getattr(model, k).copy_(v.cuda())
            r = model.solve(K, mi, me, *(qp[k].cuda() for k in ("Q", "p", "A0", "zl", "zu")), 6e-6, state=[s.cuda() for s in st])
        torch.cuda.synchronize()
        errs = {k: rel_err(getattr(r, k), getattr(ref, k)) for k in ("x", "y", "z", "xv", "H", "C")}
        print(f"B={B} n={n} m={m} h={h} K={K} wscale={scale} {mode:10s}", {k: f"{v:.1e}" for k, v in errs.items()}, flush=True)

if os.environ.get("LONG"):
    # K=100 trajectories at config-2 dimensions: every tensor-core mode against the fp32 CUDA-core path
    def long_run(B, n, mi, me, h, K, scale, seed):
        qp = {k: v.cuda() for k, v in orc.qp_instances(B, n, mi, me, seed).items()}
        prm = orc.lstm_parameters(h, K, seed, scale=scale)
        sc = ia.Scaling(n, mi + me, 10, "cuda:0")
        Q, p, A0, zl, zu = sc.scale_data(qp["Q"], qp["p"], qp["A0"], qp["zl"], qp["zu"])
        res = {}
        for mode in ("simt_fp32", "tc_3xfp16", "tc_f16f8", "tc_1xfp16"):
            model = ia.LSTM(None, 2, h, K, "cuda:0", gate_mode=mode)
            with torch.no_grad():
                for k, v in prm.items(): getattr(model, k).copy_(v.cuda())
                res[mode] = model.solve(K, mi, me, Q, p, A0, zl, zu, 6e-6, scaling=sc)
            torch.cuda.synchronize()
            if mode != "simt_fp32":
                errs = {k: rel_err(getattr(res[mode], k), getattr(res["simt_fp32"], k)) for k in ("x", "y", "z", "pri", "dual", "pri_unscaled", "dual_unscaled")}
                print(f"K={K} B={B} n={n} h={h} wscale={scale} {mode:10s} vs simt_fp32", {k: f"{v:.1e}" for k, v in errs.items()}, flush=True)
    long_run(4, 1000, 500, 500, 800, 100, 1.0, 3)
    long_run(4, 1000, 500, 500, 800, 100, 3.0, 4)
    long_run(8, 100, 50, 50, 64, 100, 1.0, 5)
    long_run(8, 100, 50, 50, 64, 100, 10.0, 6)
    long_run(4, 1000, 500, 500, 208, 100, 1.0, 7)
    sys.exit(0)
run(2, 12, 5, 7, 8, 1)
run(2, 12, 5, 7, 64, 1)
run(3, 100, 50, 50, 64, 1)
run(3, 100, 50, 50, 64, 3)
run(2, 100, 50, 50, 200, 2)
run(2, 200, 50, 50, 800, 2)
run(2, 200, 50, 50, 800, 2, scale=10.0)
run(1, 1000, 500, 500, 800, 2)
