#!/usr/bin/env python
"""Benchmark of the I-ADMM-LSTM unrolled solve path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--gate-mode M] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = the hot path over one batch of synthetic QPs: Ruiz equilibration (10 its) + K=100 unrolled
I-ADMM-LSTM iterations + the per-iteration primal/dual residual traces, for `batch` instances per GPU of
BASELINE config 2 (n=1000, 500 ineq + 500 eq, hidden_dim=800, --scaling).  The batch shards by instance
across GPUs with no collective (weak scaling: every rank solves its own `batch` instances).

Prints ONE JSON line (rank 0).  `value` = solves/s with inputs resident in HBM; `e2e` = the same through
the public API with pinned HOST buffers, H2D/D2H copies inside the timed region; `roofline` = the gate
kernel (tensor-bound) measured live with CUDA events through the library's profile hooks, plus
`roofline_kkt` for the HBM-bound KKT phase; `cpu_baseline` = the oracle port of the reference's torch
path on the host cores (bounded sample).  `--impl reference` times that CPU path alone.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "i-admm-lstm_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

N_VAR, N_INEQ, N_EQ, HIDDEN, K_ITERS, SIGMA, RUIZ_ITS = 1000, 500, 500, 800, 100, 6e-6, 10
METRIC = "QP solves/sec (K=100, n=1000, m=1000)"
UNIT = "solves/s"
WORKLOAD = ("config2: dense QP n=1000, 500 ineq + 500 eq, hidden_dim=800, --scaling (10 Ruiz its), K=100 "
            "unrolled iterations + per-iteration residual traces, random-init LSTM")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gate-mode", default="tc_f16f8", choices=["tc_3xfp16", "tc_f16f8", "tc_1xfp16", "simt_fp32"])
    ap.add_argument("--batch", type=int, default=256, help="instances per GPU")
    ap.add_argument("--iters", type=int, default=K_ITERS, help="unrolled iterations per solve (metric is quoted at 100)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-balance", action="store_true",
                    help="N > 1: keep equal shards instead of sharding the N*batch instances by each GPU's measured rate")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-batch", type=int, default=4, help="instances per CPU step of the reference arm (~0.8 s each on 16 cores)")
    ap.add_argument("--nvar", type=int, default=N_VAR, help="variables per QP (the metric is quoted at 1000; 5000 = BASELINE config 5, "
                                                             "with num_ineq = num_eq = nvar/2)")
    ap.add_argument("--hidden", type=int, default=HIDDEN, help="hidden_dim (the metric is quoted at 800; 200 is configs/QP.yaml's default)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
# synthetic inputs: generate_data.py:67-76 restated on the device (main.py:718 doubles Q on load)
# ---------------------------------------------------------------------------------------------------
def device_qp_batch(B, n, mi, me, seed, dev):
    from iadmm_b200.data import generate_qp_batch
    d = generate_qp_batch(B, n, mi, me, seed, dev)
    return d["Q"], d["p"], d["A0"], d["zl"], d["zu"]


# ---------------------------------------------------------------------------------------------------
# clocks during the timed region
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, dev):
        self.proc = None
        self.path = None
        try:
            uuid = str(torch.cuda.get_device_properties(dev).uuid)
            self.gpu_id = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
        except Exception:
            self.gpu_id = str(torch.device(dev).index or 0)

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", self.gpu_id, "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        clocks, mx, reasons, power = [], None, set(), []
        try:
            for line in open(self.path):
                f = [s.strip() for s in line.split(",")]
                if len(f) < 7:
                    continue
                try:
                    clocks.append(float(f[0])); mx = float(f[1]); power.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(self.NAMES, f[3:7]):
                    if v == "Active":
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if clocks:
            out.update(sm_mhz=statistics.median(clocks), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(clocks),
                       power_w_max=max(power) if power else None)
        return out


# ---------------------------------------------------------------------------------------------------
# the reference arm / cpu baseline: oracle port of the reference's torch CPU path
# ---------------------------------------------------------------------------------------------------
def cpu_reference_solves_per_s(steps, warmup, cpu_batch, iters):
    from oracle import iadmm_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    qp = orc.qp_instances(cpu_batch, N_VAR, N_INEQ, N_EQ, seed=17)
    prm = orc.lstm_parameters(HIDDEN, iters, seed=17)

    def one_step():
        # what main.py times: scale_data (:825-834) + the K model() calls (:881-890); residuals as in :955
        Qs, ps, As, zls, zus, sc = orc.ruiz_equilibrate(qp["Q"], qp["p"], qp["A0"], qp["zl"], qp["zu"], RUIZ_ITS)
        orc.solve(prm, iters, N_INEQ, N_EQ, Qs, ps, As, zls, zus, SIGMA, HIDDEN, scaling=sc,
                  original=(qp["Q"], qp["p"], qp["A0"]), form="dense")

    with torch.no_grad():
        for _ in range(warmup):
            one_step()
        t0 = time.perf_counter()
        for _ in range(steps):
            one_step()
        dt = time.perf_counter() - t0
    sample = (f"{steps} x ({cpu_batch} instance(s) of the same workload, Ruiz + K={iters}, dense-KKT form like "
              f"models/lstm.py:67-72), torch {torch.__version__} CPU fp32, {cores} threads")
    return cpu_batch * steps / dt, dt / steps * 1e3, cores, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    val, ms, cores, sample = cpu_reference_solves_per_s(steps, warmup, args.cpu_batch, args.iters)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_step": args.cpu_batch, "iters": args.iters},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1400.0, 1590.0, "fallback"


def ncu_traffic(kind, batch, mode, headline=True):
    """DRAM bytes per launch from the committed ncu --set full capture of this config, if any."""
    if not headline:
        return None
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return d.get(f"{kind}:{mode}:B{batch}")
        except Exception:
            return None
    return None


def run_ours(args):
    import torch.distributed as dist
    import iadmm_b200 as ia
    from ctypes import byref, c_double, c_int

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")          # keep NCCL's version banner off stdout (one JSON line only)
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world
    B, n, mi, me, h, K = args.batch, args.nvar, args.nvar // 2, args.nvar // 2, args.hidden, args.iters
    m, N = mi + me, n + mi + me
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    L = ia.lib()

    torch.manual_seed(17)                                    # identical random-init weights on every rank
    model = ia.LSTM(None, 2, h, K, dev, gate_mode=args.gate_mode).eval()
    # N > 1: the job is N*B independent instances.  GPUs of one box differ by several per cent under the power cap, so after
    # the warm-up the instances are re-sharded in proportion to each GPU's measured rate (iadmm_b200.dist.balance_by_rate: one
    # all-gather of a float per rank at set-up; the solve itself has no collective).  Every rank generates some spare instances.
    balance = world > 1 and not args.no_balance
    B_cap = B + max(8, B // 8) if balance else B
    Q, p, A0, zl, zu = device_qp_batch(B_cap, n, mi, me, 17 + rank, dev)
    full_inputs = (Q, p, A0, zl, zu)
    Q, p, A0, zl, zu = (t[:B] for t in full_inputs)
    shares = [B] * world
    scaling = ia.Scaling(n, m, RUIZ_ITS, dev)

    def hot_step():
        Qs, ps, As, zls, zus = scaling.scale_data(Q, p, A0, zl, zu)
        return model.solve(K, mi, me, Qs, ps, As, zls, zus, SIGMA, scaling=scaling)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(warmup):
            r = hot_step()
        barrier()
        if balance:
            from iadmm_b200.dist import balance_by_rate
            w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            w0.record()
            for _ in range(2):
                r = hot_step()
            w1.record()
            torch.cuda.synchronize()
            B_mine, shares = balance_by_rate(world * B, 2 * B / (w0.elapsed_time(w1) * 1e-3), cap=B_cap)
            Q, p, A0, zl, zu = (t[:B_mine] for t in full_inputs)
            r = hot_step()                                   # workspace for the new share, untimed
            B = B_mine
            barrier()
        sampler = ClockSampler(dev) if rank == 0 else None
        if sampler:
            sampler.start()
        ia._lib.check(L.iadmm_profile_begin(steps * K))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        torch.cuda.profiler.start()      # `ncu --profile-from-start off` captures exactly the timed region
        e0.record()
        for _ in range(steps):
            r = hot_step()
        e1.record()
        barrier()
        torch.cuda.profiler.stop()
        ms = e0.elapsed_time(e1)
        kkt_ms, gate_ms, tail_ms, nit = c_double(), c_double(), c_double(), c_int()
        ia._lib.check(L.iadmm_profile_end(byref(kkt_ms), byref(gate_ms), byref(tail_ms), byref(nit)))
        clocks = sampler.stop() if sampler else None
        finite = bool(torch.isfinite(r.x).all() and torch.isfinite(r.pri).all())

        # ---- e2e: same step through the public API from pinned host buffers --------------------------
        e2e = None
        if not args.no_e2e:
            host_in = [t.cpu().pin_memory() for t in (Q, p, A0, zl, zu)]
            # two device input buffers: the H2D copy of step i+1 (copy stream) overlaps the solve of step i
            dev_in = [[torch.empty_like(t) for t in (Q, p, A0, zl, zu)] for _ in range(2)]
            host_out = None
            copy_stream = torch.cuda.Stream()
            copied = [torch.cuda.Event(), torch.cuda.Event()]
            consumed = [torch.cuda.Event(), torch.cuda.Event()]

            def upload(i):
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(consumed[i % 2])          # buffer free again?
                    for d_, h_ in zip(dev_in[i % 2], host_in):
                        d_.copy_(h_, non_blocking=True)
                    copied[i % 2].record(copy_stream)

            def e2e_step(i, last):
                nonlocal host_out
                cur = torch.cuda.current_stream()
                cur.wait_event(copied[i % 2])
                Qs, ps, As, zls, zus = scaling.scale_data(*dev_in[i % 2])
                consumed[i % 2].record(cur)                          # scale_data has read the raw inputs
                if not last:
                    upload(i + 1)
                rr = model.solve(K, mi, me, Qs, ps, As, zls, zus, SIGMA, scaling=scaling)
                outs = (rr.x, rr.y, rr.z, rr.pri, rr.dual, rr.pri_unscaled, rr.dual_unscaled, rr.metrics)
                if host_out is None:
                    host_out = [torch.empty(o.shape, dtype=o.dtype).pin_memory() for o in outs]
                for h_, o in zip(host_out, outs):
                    h_.copy_(o, non_blocking=True)
                return outs

            for ev in consumed:
                ev.record(torch.cuda.current_stream())
            upload(0)
            for i in range(2):
                outs = e2e_step(i, i == 1)
            barrier()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            upload(2)                                    # every timed step's H2D copy is inside the timed region
            for i in range(2, 2 + steps):
                outs = e2e_step(i, i == 1 + steps)
            t1.record()
            barrier()
            e2e_ms = t0.elapsed_time(t1)
            h2d = sum(t.numel() * t.element_size() for t in host_in)
            d2h = sum(o.numel() * o.element_size() for o in outs)
            e2e = (e2e_ms, h2d, d2h)

    # max over ranks
    times = torch.tensor([ms, e2e[0] if e2e else 0.0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(times[0]), float(times[1])

    if rank == 0:
        hbm_gbs, tf_sus, tf_burst, peak_kind = measured_peaks()
        rows = B * N
        nit_v = max(1, nit.value)
        gate_avg_ms = gate_ms.value / nit_v
        kkt_avg_ms = kkt_ms.value / nit_v
        gate_flops = 8.0 * rows * h * h                      # logical fp32 flops of H@U (no credit for the 3-way split)
        gate_tflops = gate_flops / (gate_avg_ms * 1e-3) / 1e12 if gate_avg_ms > 0 else 0.0
        kkt_bytes = 8.0 * B * (n * n + m * n)                # Q and A0 streamed once per pass, two passes
        kkt_gbs = kkt_bytes / (kkt_avg_ms * 1e-3) / 1e9 if kkt_avg_ms > 0 else 0.0
        iter_bytes = kkt_bytes + 16.0 * rows * h + 64.0 * rows       # SURVEY section 8(d) bytes per iteration
        step_s = ms / steps * 1e-3
        hbm_frac_whole = (iter_bytes * K / step_s / 1e9) / hbm_gbs
        launches_per_step = K * 6 + 2 + (3 + 2 * RUIZ_ITS)
        line = {
            "metric": METRIC, "value": sum(shares) * steps / (ms * 1e-3), "unit": UNIT, "n_gpus": n_gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": (WORKLOAD if h == HIDDEN else WORKLOAD.replace("hidden_dim=800", "hidden_dim=%d" % h)) if n == N_VAR else
                                   "config5-style: dense QP n=%d, %d ineq + %d eq, hidden_dim=%d, --scaling, K=%d (NOT the headline workload)" % (n, mi, me, h, K),
                       "batch_per_gpu": args.batch, "instances_per_gpu": shares, "iters": K, "gate_mode": args.gate_mode,
                       "gate_arithmetic": {"tc_3xfp16": "tcgen05 fp16 hi/lo split, 3 MMAs, fp32 accumulate",
                                           "tc_f16f8": "tcgen05 fp16 MMA + 2 e4m3 correction MMAs, fp32 accumulate",
                                           "tc_1xfp16": "tcgen05 single fp16 MMA, fp32 accumulate",
                                           "simt_fp32": "fp32 FMA"}[args.gate_mode],
                       "cache": "inputs larger than L2: Q+A0 %.2f GB and LSTM state %.2f GB per GPU per iteration vs 126 MB L2"
                                % (kkt_bytes / 2 / 1e9, 16.0 * rows * h / 1e9),
                       "parallelism": ("%d x %d independent instances sharded over %d GPU(s)%s, no collective in the solve"
                                       % (n_gpus, args.batch, n_gpus, " in proportion to each GPU's measured warm-up rate" if balance else "")),
                       "results_finite": finite},
            "gpu_launches": steps * launches_per_step,
            "roofline": {"kernel": "gates_tc_pair_kernel" if args.gate_mode != "simt_fp32" else "gates_simt_kernel",
                         "bound": "tensor", "achieved": gate_tflops, "peak": tf_sus, "unit": "TFLOP/s",
                         "frac": gate_tflops / tf_sus, "traffic": ncu_traffic("gates", B, args.gate_mode, h == HIDDEN and n == N_VAR),
                         "peak_kind": "%s bf16_tflops_sustained (kernel timed inside a long step)" % peak_kind,
                         "flops_per_launch": gate_flops, "ms_per_launch": gate_avg_ms,
                         "share_of_step": gate_ms.value / ms,
                         "issued_frac": gate_tflops * {"tc_3xfp16": 3, "tc_f16f8": 2}.get(args.gate_mode, 1) / tf_sus,
                         "note": "logical fp32 flops 8*rows*h^2; the operand split issues %dx that in fp16-equivalent MMA work "
                                 "(tc_3xfp16: 3 fp16 products, tc_f16f8: 1 fp16 + 2 e4m3 at twice the rate; issued rate %.1f TFLOP/s "
                                 "= issued_frac of the peak); the kernel runs at the board power cap (see clocks), DESIGN.md section 4"
                                 % ({"tc_3xfp16": 3, "tc_f16f8": 2}.get(args.gate_mode, 1),
                                    gate_tflops * {"tc_3xfp16": 3, "tc_f16f8": 2}.get(args.gate_mode, 1))},
            "roofline_kkt": {"kernel": "kkt_pass1+combine1+pass2+combine2", "bound": "hbm", "achieved": kkt_gbs,
                             "peak": hbm_gbs, "unit": "GB/s", "frac": kkt_gbs / hbm_gbs,
                             "traffic": ncu_traffic("kkt", B, args.gate_mode, n == N_VAR),
                             "bytes_per_iteration": kkt_bytes, "ms_per_iteration": kkt_avg_ms,
                             "share_of_step": kkt_ms.value / ms, "peak_kind": "%s hbm_gbs" % peak_kind},
            "hbm_roofline_frac_whole_path": hbm_frac_whole,
            "phase_ms_per_iteration": {"kkt": kkt_avg_ms, "gates": gate_avg_ms, "tail": tail_ms.value / nit_v},
            "clocks": clocks,
        }
        if e2e:
            line["e2e"] = {"value": sum(shares) * steps / (e2e_ms * 1e-3), "unit": UNIT,
                           "h2d_bytes_per_step": e2e[1], "d2h_bytes_per_step": e2e[2], "ms_per_step": e2e_ms / steps}
        if n_gpus == 1 and not args.no_cpu_baseline:
            val, cms, cores, sample = cpu_reference_solves_per_s(3, 1, 4, K)     # 16 instances, ~12 s of CPU work; batches of 4 are the CPU path's best
            # operating point here (measured 1.3 solves/s at batch 1-4, 0.56 at batch 12: the dense KKT build falls out of cache)
            line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
