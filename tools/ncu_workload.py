"""Small workload for ncu captures: decode `n` bench.jxl frames once (plus one warm-up)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge

pkg = ge.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
name = sys.argv[2] if len(sys.argv) > 2 else "bench.jxl"
data = open(os.path.join(ROOT, "tests", "golden", name), "rb").read()
dec = pkg.BatchDecoder(0)
nch = int(sys.argv[3]) if len(sys.argv) > 3 else 4
dec.set_input([data] * n, nch, pkg.JXL_TYPE_UINT8)
for _ in range(2):
    dec.run()
    dec.wait()
print("ok", dec.stats().num_streams)
