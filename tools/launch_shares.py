"""Kernel shares of a step from an ncu launch list (`--metrics gpu__time_duration.sum --csv`): total time per kernel name
over the listed launches (serialised, cold-cache times: the shares are what compares with bench.py's per-class event times)."""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    name = r[ki].split("(")[0].replace("void ", "")
    tot[name] += float(r[vi]) / 1e6
    cnt[name] += 1
s = sum(tot.values())
print("%d launches, %.1f ms" % (sum(cnt.values()), s))
for k, v in tot.most_common():
    print("%6.1f ms  %5.1f%%  x%-4d %s" % (v, 100 * v / s, cnt[k], k))
