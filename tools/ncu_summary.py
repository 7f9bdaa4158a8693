"""Summarises an .ncu-rep (raw + source pages) into the few numbers we track per kernel."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else None  # warp-level loop iterations, for per-step figures
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print("kernel:", d.get("Kernel Name"), "grid", d.get("launch__grid_size"), "block", d.get("launch__block_size"))
    for k in ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
              "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
              "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
              "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
              "lts__t_sector_hit_rate.pct", "smsp__average_warp_latency_per_inst_issued.ratio",
              "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
              "sm__throughput.avg.pct_of_peak_sustained_elapsed"]:
        if k in d:
            print(f"  {k} = {d[k]} {units[hdr.index(k)]}")
    stalls = {k: float(v) for k, v in d.items() if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and v}
    for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:6]:
        print("  stall", k.split("issue_stalled_")[1].split("_per_issue")[0], round(v, 2))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
ops = collections.Counter()
tot = 0
for r in data:
    toks = r[ci["Source"]].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    n = int(r[ci["Instructions Executed"]])
    ops[op.split(".")[0]] += n
    tot += n
print("total warp instructions", tot, "per step" if steps else "", round(tot / steps, 1) if steps else "")
for op, c in ops.most_common(14):
    print(f"  {op:8s} {c / tot * 100:5.1f}%", round(c / steps, 1) if steps else "")
hot = sorted(data, key=lambda r: -int(r[ci["# Samples"]]))[:12]
ts = sum(int(r[ci["# Samples"]]) for r in data)
print("hottest SASS by stall samples:")
for r in hot:
    print(f"  {int(r[ci['# Samples']]) / ts * 100:5.1f}%  thr={r[ci['Avg. Threads Executed']]:>4s}  {r[ci['Source']].strip()[:90]}")
