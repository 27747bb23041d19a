"""Top source lines of a kernel by warp-stall samples, from an .ncu-rep captured with --import-source on.
usage: python tools/ncu_hot_lines.py rep.ncu-rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
data = []
for r in rows:
    if hdr is None:
        if "Source" in r and any("Samples" in c for c in r): hdr = r
        continue
    if len(r) == len(hdr): data.append(r)
if hdr is None:
    print(out[:2000]); sys.exit(1)
ci = hdr.index("Source")
si = [i for i, c in enumerate(hdr) if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)"]
si = si[0] if si else [i for i, c in enumerate(hdr) if "Samples" in c][0]
ii = [i for i, c in enumerate(hdr) if c.startswith("Instructions Executed")]
def num(x):
    try: return float(x.replace(",", ""))
    except Exception: return 0.0
tot = sum(num(r[si]) for r in data) or 1.0
print("sample column:", hdr[si], "total", tot)
for r in sorted(data, key=lambda r: -num(r[si]))[:top]:
    print("%6.2f%%  L%-5s %s" % (100 * num(r[si]) / tot, r[0], r[ci].strip()[:150]))
