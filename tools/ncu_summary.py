"""Summarise an .ncu-rep (raw page) into the handful of numbers the roofline discussion needs.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [extra-regex]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
extra = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
WANT = [r"^gpu__time_duration\.sum$", r"^dram__bytes_(read|write)\.sum$", r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$",
        r"^lts__throughput\.avg\.pct", r"^lts__t_bytes\.sum$", r"^sm__cycles_elapsed\.avg\.per_second$",
        r"^launch__(registers_per_thread|grid_size|block_size|occupancy_limit.*|waves_per_multiprocessor)$",
        r"^launch__shared_mem_per_block_dynamic$",
        r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$", r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$",
        r"^smsp__issue_active\.avg\.pct", r"^sm__inst_executed_pipe_(xu|fma|alu|lsu|uniform|tensor.*)\.avg\.pct_of_peak_sustained_active$",
        r"^sm__pipe_tensor.*cycles_active.*\.avg\.pct_of_peak_sustained_(active|elapsed)$",
        r"^sm__inst_executed_pipe_tc.*", r"^sm__mem_tensor_cycles_active\.avg\.pct_of_peak_sustained_elapsed$",
        r"^smsp__average_warps?_issue_stalled_.*_per_issue_active\.ratio$", r"^smsp__warp_issue_stalled_.*\.avg$",
        r"^l1tex__t_bytes.*\.sum$", r"^smsp__inst_executed\.sum$", r"^sm__ops_path_tensor.*\.sum$",
        r"^smsp__average_warp_latency_issue_stalled.*"]
pats = [re.compile(p) for p in WANT]
for r in rows[2:]:
    print("=== %s  (id %s)" % (r[hdr.index("Kernel Name")][:70], r[hdr.index("ID")]))
    for i, h in enumerate(hdr):
        if any(p.search(h) for p in pats) or (extra and extra.search(h)):
            v = r[i]
            if v in ("", "no data", "0"):
                continue
            print("   %-95s %-10s %s" % (h, units[i], v))
