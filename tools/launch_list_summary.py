"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.  usage: launch_list_summary.py file.csv [divisor]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]; kn = hdr.index('Kernel Name'); mv = hdr.index('Metric Value'); mu = hdr.index('Metric Unit')
agg = collections.OrderedDict(); tot = 0; cnt = 0
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    v = float(r[mv].replace(',', ''))
    if r[mu] == 'ns': v /= 1000
    elif r[mu] == 'ms': v *= 1000
    name = r[kn].split('(')[0][-60:]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; tot += v; cnt += 1
print("launches %d  total %.1f us  per unit %.1f us (%.1f launches)" % (cnt, tot, tot / div, cnt / div))
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:30]:
    print("%8.1f us/unit  %5.1f launches/unit  %7.1f us each  %s" % (t / div, c / div, t / c, k))
