"""Qualification of the exp-only tanh in the 16-warp gate kernel (development switch IADMM_TC_EPI=2, development build): worst
relative error of x^K, y^K, z^K per INSTANCE after K=100 against the fp32 CUDA-core path on the same GPU, streaming kernels, at
the headline problem size for the hidden sizes the 16-warp kernel serves.  Run once per setting of the switch.

    IADMM_B200_LIB=.../libiadmm_b200_dev.so [IADMM_TC_EPI=2] python tools/qualify_tanh.py [hidden ...]
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "i-admm-lstm_b200"))
import torch
import iadmm_b200 as ia
from bench import device_qp_batch

dev = torch.device("cuda:0")
n, K, B = 1000, 100, 16
mi = me = n // 2


def per_instance_err(a, b):
    a, b = a.double().reshape(a.shape[0], -1), b.double().reshape(b.shape[0], -1)
    return ((a - b).norm(dim=1) / b.norm(dim=1).clamp_min(1e-300))


for h in [int(v) for v in sys.argv[1:]] or [200, 384]:
    for wscale in (1.0, 2.0):
        worst = {"x": 0.0, "y": 0.0, "z": 0.0}
        for seed in (1, 2, 3):
            torch.manual_seed(seed)
            ref_model = ia.LSTM(None, 2, h, K, dev, gate_mode="simt_fp32")
            if wscale != 1.0:
                with torch.no_grad():
                    for name, prm in ref_model.named_parameters():
                        if name[0] in "WU":
                            prm.mul_(wscale)
            Q, p, A0, zl, zu = device_qp_batch(B, n, mi, me, 2000 + seed, dev)
            data = ia.Scaling(n, mi + me, 10, dev).scale_data(Q, p, A0, zl, zu)
            with torch.no_grad():
                ref = ref_model.solve(K, mi, me, *data, 6e-6, streaming=True)
                model = ia.LSTM(None, 2, h, K, dev, gate_mode="tc_f16f8")
                model.load_state_dict(ref_model.state_dict())
                r = model.solve(K, mi, me, *data, 6e-6, streaming=True)
            for k in worst:
                worst[k] = max(worst[k], float(per_instance_err(getattr(r, k), getattr(ref, k)).max()))
        print(json.dumps({"tanh": "exp-only" if os.environ.get("IADMM_TC_EPI") == "2" else "polynomial below 0.55", "hidden": h, "K": K,
                          "instances": 3 * B, "weights": "random-init x%g" % wscale,
                          "worst_instance": {k: float("%.2e" % v) for k, v in worst.items()},
                          "margin_to_1e-4": float("%.1f" % (1e-4 / max(worst.values())))}), flush=True)
