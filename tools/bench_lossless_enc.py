"""Lossless (Modular) encode throughput of the CUDA encoder: n 4K RGB8 frames per call, host buffers in, codestreams out
(the call is synchronous: H2D + kernels + host tables + emit + D2H + assembly), best of `reps`; the decode round trip of
one output is checked. Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge

pkg = ge.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
src = open(os.path.join(ROOT, "tests", "golden", "vardct_4k_natural.jxl"), "rb").read()
base = pkg.decode_batch([src], 3, np.uint8)[0]
imgs = [np.ascontiguousarray(np.roll(base, ((13 * i) % 256, (7 * i) % 256), axis=(0, 1))) for i in range(n)]
enc = pkg.JxlEncoder(lossless=True, uses_original_profile=True)
best, phases = None, None
for _ in range(reps + 1):
    t0 = time.perf_counter()
    outs = enc.encode_batch(imgs)
    dt = time.perf_counter() - t0
    if best is None or dt < best:
        best, phases = dt, enc.phase_times()
back = pkg.decode_batch([outs[0].data], 3, np.uint8)[0]
px = sum(a.shape[0] * a.shape[1] for a in imgs)
print(json.dumps({"workload": "lossless encode of %d 4K RGB8 frames per call" % n, "e2e_mpx_s": px / best / 1e6,
                  "ms_per_call": best * 1e3, "device_phases_ms": phases, "bpp": 8.0 * sum(len(o.data) for o in outs) / px,
                  "round_trip_ok": bool(np.array_equal(back, imgs[0]))}))
