"""Kernel shares of a step from an ncu launch list (`--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]
--csv`): total time per kernel name over the listed launches (serialised, cold-cache times: the SHARES are what compares
with bench.py's per-class event times). The bench command first encodes its batch on the GPU (k_enc_* / k_encl_*
launches): those are listed apart, the shares are over the decode kernels. With DRAM metrics in the list: bytes per
decode step (all decode kernels of one JxlB200DecoderRun), for profiles/traffic.json."""
import collections
import csv
import json
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
ki, mi, vi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
tot, cnt, dram = collections.Counter(), collections.Counter(), collections.Counter()
for r in rows[1:]:
    name = r[ki].split("(")[0].replace("void ", "")
    if r[mi] == "gpu__time_duration.sum":
        tot[name] += float(r[vi]) / 1e6
        cnt[name] += 1
    elif r[mi].startswith("dram__bytes"):
        dram[name] += float(r[vi])
dec = {k: v for k, v in tot.items() if not k.startswith("k_enc")}
enc = {k: v for k, v in tot.items() if k.startswith("k_enc")}
s = sum(dec.values())
runs = max(1, cnt.get("k_block_lists", 0))  # one launch per JxlB200DecoderRun of a lossy batch
print("%d decode launches in %d runs of the batch, %.1f ms serialised (%.1f ms per run)" % (sum(cnt[k] for k in dec), runs, s, s / runs))
for k, v in sorted(dec.items(), key=lambda kv: -kv[1]):
    print("%7.1f ms  %5.1f%%  x%-4d %s" % (v, 100 * v / s, cnt[k], k))
if enc:
    print("(batch preparation by the GPU encoder, not part of the step: %d launches, %.1f ms)" % (sum(cnt[k] for k in enc), sum(enc.values())))
if dram:
    b = sum(v for k, v in dram.items() if not k.startswith("k_enc")) / runs
    print("DRAM bytes per decode run, all decode kernels: %.0f" % b)
    print(json.dumps({"dram_bytes_per_step_all_kernels": b, "per_kernel": {k: v / runs for k, v in dram.items() if not k.startswith("k_enc")}}))
