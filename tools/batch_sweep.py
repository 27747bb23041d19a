"""BASELINE config 4: batch-size sweep at n=1000 (instances per GPU), solves/s and phase times per iteration.
Small batches keep Q and A0 (8 MB per instance) resident in the 126 MB L2 between the two passes and across
iterations; large batches stream them from HBM."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = []
for B in [int(b) for b in os.environ.get("BATCHES", "4,8,16,32,64,128,256,512,1024").split(",")]:
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--batch", str(B), "--steps", "2", "--warmup", "3",
                        "--no-e2e", "--no-cpu-baseline"], capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        row = {"batch": B, "solves_per_s": d["value"], "ms_per_iteration": d["phase_ms_per_iteration"],
               "kkt_GBps": d["roofline_kkt"]["achieved"], "gate_TFLOPs_logical": d["roofline"]["achieved"],
               "matrices_MB": B * 8.0, "fits_L2": B * 8.0 <= 100.0, "sm_mhz": d["clocks"]["sm_mhz"]}
    except Exception as e:
        row = {"batch": B, "error": (r.stderr or str(e))[-300:]}
    out.append(row); print(json.dumps(row), flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "batch_sweep.json"), "w"), indent=1)
