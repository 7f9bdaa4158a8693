"""How much does a batch of byte-identical frames flatter the lock-step entropy kernels? Decodes N GPU-encoded frames that
are all different (integer translations of one image) and N replicas of the first of them, one
handle alone, per-class kernel times."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge

pkg = ge.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
src = open(os.path.join(ROOT, "tests", "golden", "vardct_4k_natural.jxl"), "rb").read()
base = pkg.decode_batch([src], 3, np.uint8)[0]
enc = pkg.JxlEncoder(quality=1.0)
files = []
for i0 in range(0, n, 16):
    imgs = [np.ascontiguousarray(np.roll(base, ((13 * i) % 256, (7 * i) % 256), axis=(0, 1))) for i in range(i0, min(n, i0 + 16))]
    files += [r.data for r in enc.encode_batch(imgs, epf_iters=1)]
del enc
for name, fl in (("distinct", files), ("replicas", [files[0]] * n)):
    dec = pkg.BatchDecoder(0)
    dec.set_input(fl, 3, pkg.JXL_TYPE_UINT8)
    dec.run(); dec.wait()
    dec.set_profiling(True)
    for _ in range(3):
        dec.run(); dec.wait()
    km, r = dec.kernel_times_ex()
    print(json.dumps({"batch": name, "frames": n, "bytes": sum(len(f) for f in fl),
                      "ms": {k: round(v / r, 2) for k, v in km.items() if v > 0.05}}))
    del dec
