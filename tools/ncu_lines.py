"""Per-source-line instruction counts of one kernel from an .ncu-rep captured with --import-source on (-lineinfo build)."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, agg, src = "?", collections.Counter(), {}
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if not r[0].isdigit() or len(r) < 8:
        continue
    try:
        n = int(r[7])
    except ValueError:
        continue
    agg[(cur, int(r[0]))] += n
    src[(cur, int(r[0]))] = r[1].strip()
tot = sum(agg.values())
print("warp instructions attributed to source lines:", tot)
for k, n in agg.most_common(top):
    print("%5.1f%%  %s:%d  %s" % (100.0 * n / tot, k[0], k[1], src[k][:120]))
