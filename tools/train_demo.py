"""End-to-end sanity run of the whole path on small QPs: train the LSTM optimiser with the reference's loop
(main.py:336-358: truncated BPTT windows, Adam) through the drop-in modules, then solve UNSEEN instances with the
fused K-step solve and report residuals and the objective gap against the exact Stage-II ADMM (models/lu.py on
csrc/lu.cu) run to convergence -- north_star's "objective gap ... reported as a sanity check", with OSQP/Gurobi
replaced by the repository's own exact solver (neither is installed here).

    python tools/train_demo.py [--n 100] [--hidden 64] [--epochs 30]
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "i-admm-lstm_b200"))
import torch
import iadmm_b200 as ia
from bench import device_qp_batch

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100); ap.add_argument("--hidden", type=int, default=64)
ap.add_argument("--batch", type=int, default=32); ap.add_argument("--batches", type=int, default=8)
ap.add_argument("--epochs", type=int, default=30); ap.add_argument("--lr", type=float, default=1e-3)
ap.add_argument("--outer-t", type=int, default=100); ap.add_argument("--tl", type=int, default=50)
ap.add_argument("--gate-mode", default="tc_f16f8")
a = ap.parse_args()
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
n, mi, me, h, K, TL, B = a.n, a.n // 2, a.n // 2, a.hidden, a.outer_t, a.tl, a.batch
m = mi + me
torch.manual_seed(5)
model = ia.LSTM(None, 2, h, K, dev, gate_mode=a.gate_mode)
opt = torch.optim.Adam(model.parameters(), lr=a.lr, weight_decay=0.0)


def scaled(seed):
    Q, p, A0, zl, zu = device_qp_batch(B, n, mi, me, seed, dev)
    sc = ia.Scaling(n, m, 10, dev)
    return (Q, p, A0, zl, zu), sc, sc.scale_data(Q, p, A0, zl, zu)


def evaluate(seed):
    raw, sc, (Q, p, A0, zl, zu) = scaled(seed)
    with torch.no_grad():
        r = model.solve(K, mi, me, Q, p, A0, zl, zu, sigma=6e-6, scaling=sc, traces=True)
        # exact ADMM on the same scaled problem, to convergence (Stage II from a zero start)
        stage2 = ia.LU(dev)
        rho_vec = torch.full((B, m, 1), 0.1, device=dev); rho_vec[:, mi:] *= 1e3
        x = torch.zeros((B, n, 1), device=dev); y = torch.zeros((B, m, 1), device=dev); z = torch.zeros((B, m, 1), device=dev)
        lu = piv = A_tild = None
        for _ in range(3000):
            x, y, z, xv, A_tild, _, lu, piv = stage2(rho_vec, x, y, z, None, 6e-6, A_tild, lu, piv, Q=Q, p=p, A0=A0, zl=zl, zu=zu)
        pri_e, dual_e, _ = ia.primal_dual_loss(x, y, z, Q, p, A0)
        obj_exact = ia.obj_fn(torch.bmm(sc.D, x), Q=raw[0], p=raw[1]).squeeze()
        obj_lstm = ia.obj_fn(torch.bmm(sc.D, r.x), Q=raw[0], p=raw[1]).squeeze()
        gap = ((obj_lstm - obj_exact).abs() / obj_exact.abs().clamp_min(1e-6))
    return dict(pri=float(r.pri[-1].mean()), dual=float(r.dual[-1].mean()), pri_exact=float(pri_e.mean()), dual_exact=float(dual_e.mean()),
                obj_gap_mean=float(gap.mean()), obj_gap_max=float(gap.max()))


before = evaluate(10_000)
data = [scaled(100 + i)[2] for i in range(a.batches)]
losses = []
t_start = time.perf_counter()
for ep in range(a.epochs):
    tot = 0.0
    for (Q, p, A0, zl, zu) in data:
        x = torch.zeros((B, n, 1), device=dev); y = torch.zeros((B, m, 1), device=dev); z = torch.zeros((B, m, 1), device=dev)
        xv = torch.zeros((B, n + m, 1), device=dev); H = torch.zeros((B, n + m, h), device=dev); C = torch.zeros((B, n + m, h), device=dev)
        for w in range(K // TL):
            loss = 0.0
            for t in range(TL):
                x, y, z, xv, H, C, _, _, _ = model(t, mi, me, x, y, z, xv, 6e-6, H, C, Q=Q, p=p, A0=A0, lb=None, ub=None, zl=zl, zu=zu)
                _, _, l = ia.primal_dual_loss(x, y, z, Q, p, A0)
                loss = loss + l.mean() / K
            opt.zero_grad(); loss.backward(); opt.step()
            x, y, z, xv, H, C = (v.detach() for v in (x, y, z, xv, H, C))
            tot += float(loss.detach())
    losses.append(tot / len(data))
torch.cuda.synchronize()
train_s = time.perf_counter() - t_start
after = evaluate(10_000)
print(json.dumps({"workload": f"train {a.epochs} epochs x {a.batches} batches x {B} QPs (n={n}, {mi}+{me}, h={h}, T={K}, TL={TL}), evaluate on {B} unseen",
                  "train_seconds": train_s, "loss_first_last": [losses[0], losses[-1]], "loss_curve": losses[:: max(1, len(losses) // 10)],
                  "before": before, "after": after}))
