"""Same-box A/B of the two dense KKT pass kernels (development build, switch IADMM_KKT_TMA): load instructions
(`kkt_pass*_kernel<1,0,0>`, ld.global.nc.v4) vs bulk-copy staging (`kkt_pass*_tma_kernel`, cp.async.bulk into a 3 x 32 KB ring).
Checks that every output of a K-iteration solve is bit-identical, then times the KKT phase inside the solve through the library's
profile spans (the number bench.py reports as roofline_kkt).

    IADMM_B200_LIB=.../libiadmm_b200_dev.so python tools/kkt_tma_ab.py [IADMM_KKT_TMA | IADMM_PDL]
"""
import json
import os
import sys
from ctypes import byref, c_double, c_int

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "i-admm-lstm_b200"))
import torch
import iadmm_b200 as ia
from bench import device_qp_batch, SIGMA

dev = torch.device("cuda", 0)
SWITCH = sys.argv[1] if len(sys.argv) > 1 else "IADMM_KKT_TMA"      # or IADMM_PDL: programmatic dependent launches in the KKT phase


def run(tma, B, n, mi, me, h, K, seed, time_it):
    os.environ[SWITCH] = "1" if tma else "0"
    torch.manual_seed(seed)
    model = ia.LSTM(None, 2, h, K, dev).eval()
    Q, p, A0, zl, zu = device_qp_batch(B, n, mi, me, seed, dev)
    sc = ia.Scaling(n, mi + me, 10, dev)
    L = ia.lib()
    with torch.no_grad():
        Qs, ps, As, zls, zus = sc.scale_data(Q, p, A0, zl, zu)
        r = model.solve(K, mi, me, Qs, ps, As, zls, zus, SIGMA, scaling=sc, streaming=True)
        torch.cuda.synchronize()
        ms = None
        if time_it:
            ia._lib.check(L.iadmm_profile_begin(3 * K))
            for _ in range(3):
                r = model.solve(K, mi, me, Qs, ps, As, zls, zus, SIGMA, scaling=sc, streaming=True)
            torch.cuda.synchronize()
            k, g, t, nit = c_double(), c_double(), c_double(), c_int()
            ia._lib.check(L.iadmm_profile_end(byref(k), byref(g), byref(t), byref(nit)))
            ms = k.value / max(1, nit.value)
    return r, ms


def main():
    shapes = [(256, 1000, 500, 500, 200, 20, True), (256, 1000, 500, 500, 800, 10, True), (24, 5000, 2500, 2500, 64, 6, True),
              (3, 260, 70, 54, 64, 5, False), (2, 1028, 37, 501, 64, 4, False), (5, 96, 40, 24, 64, 4, False), (2, 2052, 0, 0, 64, 3, False)]
    for B, n, mi, me, h, K, time_it in shapes:
        a, ta = run(False, B, n, mi, me, h, K, 7, time_it)
        b, tb = run(True, B, n, mi, me, h, K, 7, time_it)
        same = all(torch.equal(getattr(a, k), getattr(b, k)) for k in ("x", "y", "z", "xv", "H", "C", "pri", "dual", "pri_unscaled", "metrics"))
        byts = 8.0 * B * (n * n + (mi + me) * n)
        row = {"B": B, "n": n, "m": mi + me, "hidden_dim": h, "K": K, "bit_identical": same}
        if time_it:
            row.update(switch=SWITCH, kkt_ms_off=round(ta, 4), kkt_ms_on=round(tb, 4), tbs_off=round(byts / ta / 1e9, 3), tbs_on=round(byts / tb / 1e9, 3))
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
