"""Prints the headline numbers of a bench.py JSON line."""
import json
import sys

j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(j["config"].get("batch_per_gpu"), "frames", "| value %.1f %s | e2e %.1f | ms/step %.1f" %
      (j["value"], j["unit"], j["e2e"]["value"], j["ms_per_step"]), "| ok", j["config"].get("golden_checksum_ok"))
print("  kernels ms:", {k: round(v, 2) for k, v in j["roofline"]["all_kernels_ms"].items() if v > 0.01})
print("  roofline:", j["roofline"]["kernel"], "achieved %.2f GB/s frac %.5f" % (j["roofline"]["achieved"], j["roofline"]["frac"]))
if "cpu_baseline" in j:
    print("  cpu:", j["cpu_baseline"])
