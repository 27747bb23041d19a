"""BASELINE config 1 (the reference's own CPU-runnable case): n=100, 50+50, h=64, batch 64, K=100, random-init LSTM.
Launch-latency bound on a GPU (the whole batch is 12 MB): eager launches vs one CUDA-graph replay."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "i-admm-lstm_b200"))
import torch
import iadmm_b200 as ia
from oracle import iadmm_oracle as orc
dev = "cuda:0"
B, n, mi, me, h, K = 64, 100, 50, 50, 64, 100
qp_cpu = orc.qp_instances(B, n, mi, me, seed=17)
prm = orc.lstm_parameters(h, K, seed=17)
qp = {k: v.to(dev) for k, v in qp_cpu.items()}
out = {"workload": "config1: n=100, 50+50, h=64, batch 64, K=100, no scaling"}
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for mode, streaming in (("tc_f16f8", False), ("tc_f16f8", True), ("simt_fp32", True)):
    model = ia.LSTM(None, 2, h, K, dev, gate_mode=mode)
    with torch.no_grad():
        for k, v in prm.items(): getattr(model, k).copy_(v.to(dev))
        args = (K, mi, me, qp["Q"], qp["p"], qp["A0"], qp["zl"], qp["zu"], 6e-6)
        kw = dict(streaming=streaming)
        ms = timeit(lambda: model.solve(*args, **kw))
        m = mi + me
        st = [torch.zeros((B, n, 1), device=dev), torch.zeros((B, m, 1), device=dev), torch.zeros((B, m, 1), device=dev),
              torch.zeros((B, n + m, 1), device=dev), torch.zeros((B, n + m, h), device=dev), torch.zeros((B, n + m, h), device=dev)]
        work = [s.clone() for s in st]
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            model.solve(*args, state=work, inplace=True, **kw)
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        for w, s in zip(work, st): w.copy_(s)
        with torch.cuda.graph(g):
            res = model.solve(*args, state=work, inplace=True, **kw)
        def replay():
            for w, s in zip(work, st): w.copy_(s)
            g.replay()
        ms_g = timeit(replay)
    name = mode + ("_streaming" if streaming and mode != "simt_fp32" else "_resident" if not streaming else "")
    out[name] = {"eager_ms": ms, "eager_solves_per_s": B / ms * 1e3, "graph_ms": ms_g, "graph_solves_per_s": B / ms_g * 1e3}
torch.set_num_threads(os.cpu_count())
with torch.no_grad():
    orc.solve(prm, 3, mi, me, qp_cpu["Q"], qp_cpu["p"], qp_cpu["A0"], qp_cpu["zl"], qp_cpu["zu"], 6e-6, h)
    t0 = time.perf_counter()
    orc.solve(prm, K, mi, me, qp_cpu["Q"], qp_cpu["p"], qp_cpu["A0"], qp_cpu["zl"], qp_cpu["zu"], 6e-6, h, form="dense")
    dt = time.perf_counter() - t0
out["cpu_oracle_port"] = {"s": dt, "solves_per_s": B / dt, "cores": os.cpu_count()}
print(json.dumps(out))
