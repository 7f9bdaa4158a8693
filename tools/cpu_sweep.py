"""Randomised CPU sweep: the kernels' device functions compiled for the host (tests/emul) against the oracle, on
seeded random sizes / options (decode: all 27 strategies, random side information, 0-3 EPF iterations, 1-3 passes, four
output formats; encode: distances 0.3 ... 12, with / without Gaborish). Longer than the test suite wants to be:

    python tools/cpu_sweep.py [seconds per direction, default 120]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import emul_lib  # noqa: E402
import jxlo  # noqa: E402
import vardct_cases as vc  # noqa: E402

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(12345)
t0, n, bad = time.time(), 0, 0
while time.time() - t0 < seconds:
    h, w = int(rng.integers(1, 330)), int(rng.integers(1, 330))
    img = vc.crop(h, w, int(rng.integers(0, 1000)), int(rng.integers(0, 1700)))
    kw = dict(strategy_mode=int(rng.choice([0, 1, 1, 2, 2])), gab=bool(rng.integers(2)), epf_iters=int(rng.integers(0, 4)),
              seed=int(rng.integers(1, 1 << 20)), random_side_info=bool(rng.integers(2)),
              distance=float(rng.choice([0.5, 1.0, 2.0, 4.0])), num_passes=int(rng.choice([1, 1, 1, 2, 3])))
    data = jxlo.encode_vardct(img, **kw)
    nc, dt = [(3, jxlo.UINT8), (4, jxlo.UINT16), (3, jxlo.FLOAT), (4, jxlo.UINT8)][int(rng.integers(4))]
    got = emul_lib.decode([data], nc, dt, [(h, w)])[0]
    n += 1
    if not np.array_equal(got.view(np.uint8), jxlo.decode(data, nc, dt).view(np.uint8)):
        bad += 1
        print("decode MISMATCH", h, w, kw, nc, dt)
print("decode: %d cases, %d mismatches" % (n, bad))
rng = np.random.default_rng(777)
t0, n, bad2 = time.time(), 0, 0
while time.time() - t0 < seconds:
    h, w = int(rng.integers(1, 400)), int(rng.integers(1, 400))
    img = vc.synthetic(h, w, int(rng.integers(1, 1000))) if rng.integers(3) == 0 else vc.crop(h, w, int(rng.integers(0, 1000)), int(rng.integers(0, 1700)))
    kw = dict(strategy_mode=int(rng.choice([0, 2, 2])), gab=bool(rng.integers(2)), epf_iters=int(rng.integers(0, 4)),
              distance=float(rng.choice([0.3, 0.5, 1.0, 1.6, 2.0, 3.0, 6.0, 12.0])), dc_smoothing=bool(rng.integers(2)))
    n += 1
    if emul_lib.encode(img, **kw) != jxlo.encode_vardct(img, dc_tree=1, **kw):
        bad2 += 1
        print("encode MISMATCH", h, w, kw)
print("encode: %d cases, %d mismatches" % (n, bad2))
sys.exit(1 if bad or bad2 else 0)
