"""How many of the three fp16 hi/lo products do the two backward GEMMs of a training window need?  (development build only)

The backward of one iteration runs H_bar = D U^T (GEMM 1: the adjoint that travels back through the window) and
U_bar = H^T D (GEMM 2: the weight gradient, contraction over all B*N coordinate rows), each as A_lo B_hi + A_hi B_lo +
A_hi B_hi on tcgen05.  The switches IADMM_GEMM1_MASK / IADMM_GEMM2_MASK of the development library drop products
(1 = A_lo B_hi, 2 = A_hi B_lo, 4 = A_hi B_hi).  For every mask pair this script runs ONE full truncated-BPTT window
(TL iterations, config-3 shape by default) and reports the relative error of every parameter gradient against the run
with fp32 CUDA-core GEMMs in the backward (IADMM_TRAIN_SIMT_GEMM=1) on the same inputs and weights.

    IADMM_B200_LIB=.../libiadmm_b200_dev.so python tools/gemm_mask_study.py [batch TL weight_scale]
"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "i-admm-lstm_b200"))
import torch
from bench import device_qp_batch, RUIZ_ITS, SIGMA
import iadmm_b200 as ia


def window_grads(model, scaled, B, n, mi, me, h, TL, dev):
    m, N = mi + me, n + mi + me
    for p in model.parameters():
        p.grad = None
    state = (torch.zeros((B, n, 1), device=dev), torch.zeros((B, m, 1), device=dev), torch.zeros((B, m, 1), device=dev),
             torch.zeros((B, N, 1), device=dev), torch.zeros((B, N, h), device=dev), torch.zeros((B, N, h), device=dev))
    loss, _ = model.train_window(TL, mi, me, *scaled, SIGMA, state, loss_scale=1.0 / TL, inplace=True)
    torch.cuda.synchronize()
    return float(loss), {k: p.grad.detach().double().clone() for k, p in model.named_parameters() if p.grad is not None}


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    TL = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    wscale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    n, mi, me, h = 1000, 500, 500, 800
    dev = torch.device("cuda", 0)
    torch.manual_seed(17)
    model = ia.LSTM(None, 2, h, TL, dev)
    if wscale != 1.0:
        with torch.no_grad():
            for k, p in model.named_parameters():
                if k[:2] in ("W_", "U_"):
                    p.mul_(wscale)
        model.invalidate_packed()
    raw = device_qp_batch(B, n, mi, me, 17, dev)
    scaled = ia.Scaling(n, mi + me, RUIZ_ITS, dev).scale_data(*raw)
    os.environ["IADMM_TRAIN_SIMT_GEMM"] = "1"
    ref_loss, ref = window_grads(model, scaled, B, n, mi, me, h, TL, dev)
    os.environ["IADMM_TRAIN_SIMT_GEMM"] = "0"
    rows = []
    for m1, m2 in ((7, 7), (6, 7), (5, 7), (4, 7), (7, 6), (7, 5), (7, 4), (6, 6), (6, 4), (4, 4)):
        os.environ["IADMM_GEMM1_MASK"], os.environ["IADMM_GEMM2_MASK"] = str(m1), str(m2)
        loss, g = window_grads(model, scaled, B, n, mi, me, h, TL, dev)
        err = {k: float((g[k] - ref[k]).norm() / ref[k].norm()) for k in ref if float(ref[k].norm()) > 0}
        grp = {"U": max(v for k, v in err.items() if k.startswith("U_")), "W": max(v for k, v in err.items() if k.startswith("W_")),
               "b": max(v for k, v in err.items() if k.startswith("b_")), "rho_alpha": max(err.get("rho", 0.0), err.get("alpha", 0.0))}
        row = {"batch": B, "TL": TL, "weight_scale": wscale, "gemm1_mask": m1, "gemm2_mask": m2, "loss": loss, "ref_loss": ref_loss,
               "worst_rel_err_by_group": {k: float("%.2e" % v) for k, v in grp.items()}}
        rows.append(row); print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
