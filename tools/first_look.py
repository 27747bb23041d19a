"""Quick device timings of the individual phases at config-2 shape (development aid, not the bench)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "i-admm-lstm_b200"))
import torch
import iadmm_b200 as ia

dev = "cuda:0"
B = int(os.environ.get("B", 64)); n = 1000; mi = me = 500; h = int(os.environ.get("H", 800)); K = 4
g = torch.Generator(device=dev).manual_seed(1)
Q = torch.diag_embed(torch.rand((B, n), device=dev, generator=g))
A0 = torch.randn((B, mi + me, n), device=dev, generator=g)
p = torch.rand((B, n, 1), device=dev, generator=g)
zl = -torch.rand((B, mi + me, 1), device=dev, generator=g); zu = torch.rand((B, mi + me, 1), device=dev, generator=g)
x = torch.randn((B, n, 1), device=dev, generator=g); y = torch.randn((B, mi + me, 1), device=dev, generator=g); z = torch.randn((B, mi + me, 1), device=dev, generator=g)

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

ms = timeit(lambda: ia.primal_dual_loss(x, y, z, Q, p, A0))
print(f"primal_dual_loss (1 pass over Q,A0): {ms:.3f} ms  -> {B*4*(n*n+(mi+me)*n)/ms/1e6:.0f} GB/s")
sc = ia.Scaling(n, mi + me, 10, dev)
ms = timeit(lambda: sc.scale_data(Q, p, A0, zl, zu), reps=3)
print(f"ruiz 10 its: {ms:.3f} ms -> {B*4*(n*n+(mi+me)*n)*21.5/ms/1e6:.0f} GB/s of (1+2*10+.5) passes")
for mode in os.environ.get("MODES", "simt_fp32").split(","):
    model = ia.LSTM(None, 2, h, 100, dev, gate_mode=mode)
    with torch.no_grad():
        ms1 = timeit(lambda: model.solve(1, mi, me, Q, p, A0, zl, zu, 6e-6, traces=False), reps=2, warm=1)
        ms5 = timeit(lambda: model.solve(5, mi, me, Q, p, A0, zl, zu, 6e-6, traces=False), reps=2, warm=1)
    it = (ms5 - ms1) / 4
    print(f"{mode}: {it:.3f} ms / iteration at B={B} h={h} -> {B/(it*100/1e3):.1f} solves/s (K=100), "
          f"gate flops {8*B*(n+mi+me)*h*h/it/1e9:.1f} TFLOP/s (incl. KKT time)")
