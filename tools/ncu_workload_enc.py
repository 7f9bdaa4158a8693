"""Workload for ncu captures on bench-shaped streams: n different 4K frames written by this repo's GPU encoder (as
bench.py's decode batch), decoded twice."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge

pkg = ge.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
src = open(os.path.join(ROOT, "tests", "golden", "vardct_4k_natural.jxl"), "rb").read()
base = pkg.decode_batch([src], 3, np.uint8)[0]
enc = pkg.JxlEncoder(quality=1.0)
imgs = [np.ascontiguousarray(np.roll(base, ((13 * i) % 256, (7 * i) % 256), axis=(0, 1))) for i in range(n)]
files = [r.data for r in enc.encode_batch(imgs, epf_iters=1)]
del enc
dec = pkg.BatchDecoder(0)
dec.set_input(files, 3, pkg.JXL_TYPE_UINT8)
for _ in range(2):
    dec.run()
    dec.wait()
print("ok", dec.stats().num_streams)
