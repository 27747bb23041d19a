"""The reference's LITERAL test loop (main.py:837-843, 874-890, 955) through the drop-in modules -- what a user gets who only
swaps the imports: scale_data, zero state, then K times { model(t, ...) ; primal_dual_loss(...) } with a re-bind of the state,
against the fused `model.solve(K, ...)` of the same work.

    python tools/dropin_loop.py [batch hidden K]
"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "i-admm-lstm_b200"))
import torch
from bench import device_qp_batch, RUIZ_ITS, SIGMA
import iadmm_b200 as ia


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    h = int(sys.argv[2]) if len(sys.argv) > 2 else 800
    K = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    n, mi, me = 1000, 500, 500
    m, N = mi + me, n + mi + me
    dev = torch.device("cuda", 0)
    torch.manual_seed(17)
    model = ia.LSTM(None, 2, h, K, dev).eval()
    raw = device_qp_batch(B, n, mi, me, 17, dev)
    scaling = ia.Scaling(n, m, RUIZ_ITS, dev)

    def loop(with_loss, kkt=True):
        model.materialize_kkt = kkt
        Q, p, A0, zl, zu = scaling.scale_data(*raw)
        x = torch.zeros((B, n, 1), device=dev); y = torch.zeros((B, m, 1), device=dev); z = torch.zeros((B, m, 1), device=dev)
        xv = torch.zeros((B, N, 1), device=dev); H = torch.zeros((B, N, h), device=dev); C = torch.zeros((B, N, h), device=dev)
        tot = None
        for t in range(K):
            x, y, z, xv, H, C, _, _, _ = model(t, mi, me, x, y, z, xv, SIGMA, H, C, Q=Q, p=p, A0=A0, lb=None, ub=None, zl=zl, zu=zu)
            if with_loss:
                pri, dual, tot = ia.primal_dual_loss(x, y, z, Q, p, A0)
        return x, tot

    def fused():
        Q, p, A0, zl, zu = scaling.scale_data(*raw)
        return model.solve(K, mi, me, Q, p, A0, zl, zu, SIGMA, scaling=scaling).x, None

    out = {}
    with torch.no_grad():
        for name, fn in (("fused_solve", fused), ("forward_loop", lambda: loop(False)), ("forward_loop_with_primal_dual_loss", lambda: loop(True)),
                         ("forward_loop_shared_kkt", lambda: loop(False, "shared")), ("forward_loop_no_kkt", lambda: loop(False, False))):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(2):
                xk, _ = fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 2
            out[name] = {"ms_per_solve_batch": round(ms, 2), "solves_per_s": round(B / (ms * 1e-3), 1)}
            out[name]["x_checksum"] = float(xk.double().abs().sum())
    print(json.dumps({"batch": B, "hidden_dim": h, "K": K, "resumed_calls": model.resumed_calls, **out}))


if __name__ == "__main__":
    main()
