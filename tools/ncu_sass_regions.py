"""Where a kernel's instructions and stall samples go, by SASS address range (from an .ncu-rep with source info).
usage: python tools/ncu_sass_regions.py rep.ncu-rep [bucket]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 300
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; data = rows[2:]
ie = hdr.index("Instructions Executed"); ss = hdr.index("# Samples"); src = hdr.index("Source")
stall_cols = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(float(r[ie]) for r in data); tots = sum(float(r[ss]) for r in data)
print("total warp-instructions %.0f, samples %.0f, SASS lines %d" % (tot, tots, len(data)))
for i in range(0, len(data), B):
    blk = data[i:i + B]
    e = sum(float(r[ie]) for r in blk); s = sum(float(r[ss]) for r in blk)
    if e / tot < 0.003 and s / tots < 0.003: continue
    ops = {}
    for r in blk:
        t = r[src].split()
        op = t[1] if t[0].startswith('@') else t[0]
        ops[op] = ops.get(op, 0) + float(r[ie])
    st = {}
    for c in stall_cols:
        st[hdr[c]] = sum(float(r[c] or 0) for r in blk)
    top = sorted(ops.items(), key=lambda x: -x[1])[:5]
    tst = sorted(st.items(), key=lambda x: -x[1])[:3]
    print("%5d-%5d inst %5.1f%% samp %5.1f%% | %s | %s" % (i, i + B, 100 * e / tot, 100 * s / tots,
          " ".join("%s:%.0f%%" % (k, 100 * v / max(e, 1)) for k, v in top), " ".join("%s:%.0f%%" % (k[6:], 100 * v / max(s, 1)) for k, v in tst)))
