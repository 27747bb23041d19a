"""Development aid: the gate-kernel variants (single CTA / CTA pair / 4-CTA multicast cluster) must give
bit-identical iterates (per-row arithmetic does not depend on the tiling)."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "i-admm-lstm_b200"))
    import torch, hashlib
    import iadmm_b200 as ia
    from bench import device_qp_batch
    dev = "cuda:0"
    out = {}
    for (B, n, h, K) in ((5, 1000, 800, 6), (3, 200, 208, 5), (9, 100, 64, 7)):
        torch.manual_seed(3)
        for mode in ("tc_3xfp16", "tc_f16f8", "tc_1xfp16"):
            model = ia.LSTM(None, 2, h, K, dev, gate_mode=mode)
            Q, p, A0, zl, zu = device_qp_batch(B, n, n // 2, n // 2, 5, dev)
            with torch.no_grad():
                r = model.solve(K, n // 2, n // 2, Q, p, A0, zl, zu, 6e-6)
            torch.cuda.synchronize()
            blob = b"".join(t.cpu().numpy().tobytes() for t in (r.x, r.y, r.z, r.xv, r.H, r.C, r.pri, r.dual))
            out[f"{B}x{n}x{h}:{mode}"] = hashlib.sha1(blob).hexdigest()[:16]
    print(json.dumps(out))
    sys.exit(0)
res = {}
for name, env in (("quad", {"IADMM_TC_QUAD": "1"}), ("pair", {"IADMM_TC_QUAD": "0"})):
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, __file__, "child"], env=e, capture_output=True, text=True, timeout=600)
    try:
        res[name] = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        print(name, "FAILED", r.stderr[-600:]); res[name] = {}
keys = sorted(res["pair"])
ok = True
for k in keys:
    row = {v: res[v].get(k) for v in res}
    same = len({h_ for v, h_ in row.items() if h_ is not None}) == 1
    ok &= same
    print(("OK  " if same else "DIFF"), k, row)
print("ALL IDENTICAL" if ok else "MISMATCH")
