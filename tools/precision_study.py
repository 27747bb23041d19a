"""CPU-only study (no GPU needed): which operand format of the gate product H @ U keeps the K=100 iterates within
north_star's 1e-4?  The oracle's LSTM cell is re-run with the product emulated as
    fp16(H 2^14) fp16(U 2^s)  +  Q(res_H) Q(U)  +  Q(H) Q(res_U)
for several choices of the correction format Q: none (single fp16 product), e4m3 with fixed power-of-two scalings (the
shipped F16F8 mode), and e2m1 (fp4) with a power-of-two scale per block of 16 / 32 K elements (what a block-scaled
tcgen05 kind::mxf4nvf4 / kind::mxf4 correction would compute).  Everything else is the fp32 oracle; errors are against
the plain fp32 oracle run on the same inputs.

    python tools/precision_study.py [n hidden K seeds...]
"""
import os, sys, json, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import iadmm_oracle as orc

E2M1 = torch.tensor([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0])

def q_e4m3(x):
    return x.to(torch.float8_e4m3fn).float()

def q_e2m1_block(x, block):
    """e2m1 with one power-of-two scale per `block` consecutive elements of the last dim (ue8m0 block scaling)."""
    shp = x.shape
    k = shp[-1]
    pad = (-k) % block
    if pad:
        x = torch.nn.functional.pad(x, (0, pad))
    xb = x.reshape(*x.shape[:-1], -1, block)
    mx = xb.abs().amax(dim=-1, keepdim=True)
    scale = torch.exp2(torch.ceil(torch.log2(torch.clamp(mx, min=1e-30) / 6.0)))
    y = xb / scale
    mag = y.abs().clamp(max=6.0)
    idx = torch.bucketize(mag, (E2M1[1:] + E2M1[:-1]) / 2)          # nearest representable magnitude
    q = E2M1[idx] * torch.sign(y) * scale
    q = q.reshape(*x.shape)
    return q[..., :k] if pad else q

def fp16_neighbours(x):
    """(down, up) fp16 neighbours of x (fp32 tensor of finite values in fp16's normal range) and the position of x between them"""
    r = x.half().float()
    up_of_r = torch.nextafter(r.half(), torch.tensor(float("inf")).half()).float()
    dn_of_r = torch.nextafter(r.half(), torch.tensor(float("-inf")).half()).float()
    dn = torch.where(r <= x, r, dn_of_r)
    up = torch.where(r <= x, up_of_r, r)
    up = torch.where(dn == x, dn, up)                       # exactly representable: no rounding at all
    f = torch.where(up > dn, (x - dn) / (up - dn).clamp(min=1e-30), torch.zeros_like(x))
    return dn, up, f

_DITHER = {}
def dithered_uh(Us, V, it):
    """variant `it % V` of a V-way stratified dithered fp16 rounding of Us: element e rounds up in round(f_e V) of the V
    variants (a per-element random phase decides which), so the average over V consecutive iterations is within ulp/(2V)."""
    key = (Us.data_ptr(), V)
    if key not in _DITHER:
        dn, up, f = fp16_neighbours(Us)
        g = torch.Generator().manual_seed(1234)
        phase = torch.randint(0, V, Us.shape, generator=g)
        _DITHER.clear(); _DITHER[key] = (dn, up, f, phase)
    dn, up, f, phase = _DITHER[key]
    thr = (((it + phase) % V).float() + 0.5) / V
    return torch.where(f > thr, up, dn)

def make_cell(mode):
    counter = {"it": 0}
    def cell(prm, feats, H, C):
        Ucat = torch.cat([prm[f"U_{g}"] for g in orc.GATES], dim=1)          # [h, 4h]
        mxu = float(Ucat.abs().max())
        us = 2.0 ** (12 - math.frexp(mxu)[1] + 1) if mxu > 0 else 1.0        # max|U| us in [2^12, 2^13)
        if mode.startswith("i8x"):
            # 16-bit FIXED-point operands split into bytes for tcgen05 kind::i8 (s32 accumulation, emulated exactly in float64):
            # H16 = round(H 2^15), U16 = round(U su) with max|U| su in [2^14, 2^15); x = 256 hi + lo, hi signed, lo in [0, 255].
            # i8x3 = hi hi + hi lo + lo hi (1.5 MMA units at the fp8 rate), i8x4 adds lo lo (the exact 16-bit product, 2 units).
            su16 = 2.0 ** (15 - math.frexp(mxu)[1]) if mxu > 0 else 1.0
            H16 = torch.round(H.double() * 32768.0).clamp(-32768, 32767)
            U16 = torch.round(Ucat.double() * su16).clamp(-32768, 32767)
            Hhi = torch.floor(H16 / 256.0); Hlo = H16 - 256.0 * Hhi
            Uhi = torch.floor(U16 / 256.0); Ulo = U16 - 256.0 * Uhi
            acc1 = Hhi @ Uhi
            acc2 = Hhi @ Ulo + Hlo @ Uhi
            prod = (acc1.float() * 65536.0) + (acc2.float() * 256.0)           # int32 -> fp32 in the epilogue
            if mode == "i8x4":
                prod = prod + (Hlo @ Ulo).float()
            HU = prod / (32768.0 * su16)
            h = H.shape[-1]
            pre = {g: feats @ prm[f"W_{g}"] + HU[..., i * h:(i + 1) * h] + prm[f"b_{g}"] for i, g in enumerate(orc.GATES)}
            gate_i = torch.sigmoid(pre["i"]); gate_f = torch.sigmoid(pre["f"]); gate_o = torch.sigmoid(pre["o"])
            cand = torch.tanh(pre["u"])
            C = gate_i * cand + gate_f * C
            H = gate_o * torch.tanh(C)
            return H, C, H @ prm["W_h"] + prm["b_h"]
        Hs = H * 16384.0
        if "_fb" in mode:                # temporal error feedback on the fp16 image of H (first-order noise shaping)
            carry = counter.get("carry")
            tgt = Hs if carry is None or carry.shape != Hs.shape else Hs + carry
            Hh = tgt.half().float()
            counter["carry"] = q_e4m3((tgt - Hh) * 32.0) / 32.0       # the residual is kept as the e4m3 plane
            Rh = Hs - Hh
        else:
            Hh = Hs.half().float(); Rh = Hs - Hh
        Us = Ucat * us
        Uh = Us.half().float(); Ru = Us - Uh
        if "_d" in mode:                 # dithered rounding of U, a different variant every iteration
            V = int(mode.split("_d")[1])
            Uh = dithered_uh(Us, V, counter["it"]); counter["it"] += 1
            Ru = Us - Uh
        prod = Hh @ Uh
        if mode.startswith("f16+T1"):
            prod = prod + (q_e4m3(Rh * 32.0) @ q_e4m3(Uh / 32.0))
        elif mode == "f16f8":
            prod = prod + (q_e4m3(Rh * 32.0) @ q_e4m3(Uh / 32.0)) + (q_e4m3(Hs / 64.0) @ q_e4m3(Ru * 64.0))
        elif mode.startswith("fp4b"):
            blk = int(mode[4:])
            # K is the last dim of H and the FIRST dim of U: block along K for both operands
            qU = lambda M: q_e2m1_block(M.t().contiguous(), blk).t()
            prod = prod + (q_e2m1_block(Rh, blk) @ qU(Uh)) + (q_e2m1_block(Hs, blk) @ qU(Ru))
        elif mode.startswith("f16+T2"):          # only the correction for the rounding of U (a fixed perturbation of the model)
            prod = prod + (q_e4m3(Hs / 64.0) @ q_e4m3(Ru * 64.0))
        elif mode == "f16+exact":
            prod = prod + Rh @ Uh + Hh @ Ru
        HU = prod / (16384.0 * us)
        h = H.shape[-1]
        pre = {g: feats @ prm[f"W_{g}"] + HU[..., i * h:(i + 1) * h] + prm[f"b_{g}"] for i, g in enumerate(orc.GATES)}
        gate_i = torch.sigmoid(pre["i"]); gate_f = torch.sigmoid(pre["f"]); gate_o = torch.sigmoid(pre["o"])
        cand = torch.tanh(pre["u"])
        C = gate_i * cand + gate_f * C
        H = gate_o * torch.tanh(C)
        return H, C, H @ prm["W_h"] + prm["b_h"]
    return cell

def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    h = int(sys.argv[2]) if len(sys.argv) > 2 else 800
    K = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    seeds = [int(s) for s in sys.argv[4:]] or [41, 42, 43]
    torch.set_num_threads(os.cpu_count())
    plain = orc.lstm_cell
    rows = []
    for seed in seeds:
        qp = orc.qp_instances(1, n, n // 2, n // 2, seed=seed)
        prm = orc.lstm_parameters(h, K, seed=seed)
        Qs, ps, As, zls, zus, so = orc.ruiz_equilibrate(qp["Q"], qp["p"], qp["A0"], qp["zl"], qp["zu"], 10)
        with torch.no_grad():
            orc.lstm_cell = plain
            ref = orc.solve(prm, K, n // 2, n // 2, Qs, ps, As, zls, zus, 6e-6, h, form="block")
            for mode in os.environ.get("MODES", "fp16x1,f16f8,fp4b16,fp4b32,f16+exact").split(","):
                orc.lstm_cell = make_cell(mode)
                r = orc.solve(prm, K, n // 2, n // 2, Qs, ps, As, zls, zus, 6e-6, h, form="block")
                row = {"seed": seed, "mode": mode, **{k: float("%.2e" % rel(getattr(r, k), getattr(ref, k))) for k in ("x", "y", "z", "pri", "dual")}}
                rows.append(row); print(json.dumps(row), flush=True)
        orc.lstm_cell = plain
    if not os.environ.get("MODES"):
      json.dump({"what": __doc__.split("\n\n")[0], "n": n, "hidden": h, "K": K, "rows": rows},
              open(os.path.join(ROOT, "profiles", "r01_precision_study_cpu.json"), "w"), indent=1)
