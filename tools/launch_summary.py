"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: python tools/launch_summary.py file.csv [steps]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]
ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split('(')[0][:60]
    v = float(r[vi].replace(',', ''))
    v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-62s n=%4d total=%9.1f us  avg=%8.1f us  share=%5.1f%%" % (k, n, t, t / n, 100 * t / tot))
print("total %.1f us over %d steps = %.1f us/step (ncu: cold caches, serialised launches -- compare shares)" % (tot, steps, tot / steps))
