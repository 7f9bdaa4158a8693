"""The -O3 -march=x86-64-v3 build of the oracle (oracle/libjxlo_fast.so, used by bench.py's CPU timing arms only) gives
the same bytes as the -O2 checker build: decode of the fixtures and of a lossy frame, and the encoder's output."""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

from conftest import GOLDEN, ROOT, read_golden

CHILD = r"""
import hashlib, json, sys
sys.path.insert(0, %r)
import numpy as np
import jxlo, vardct_cases as vc
if sys.argv[1] == "fast":
    jxlo.use_fast_build()
out = {}
for name, nc, dt in [("sample.jxl", 4, jxlo.UINT16), ("sample_jpg.jxl", 3, jxlo.UINT8), ("sample_grey.jxl", 1, jxlo.UINT16),
                     ("sample_grey.jxl", 3, jxlo.FLOAT)]:
    data = open(%r + "/" + name, "rb").read()
    out[name + str(nc) + str(dt)] = hashlib.sha256(jxlo.decode(data, nc, dt).tobytes()).hexdigest()
img = vc.crop(300, 400, 100, 200)
for kw in (dict(strategy_mode=1, random_side_info=True, seed=3, epf_iters=3), dict(strategy_mode=2, distance=2.0)):
    enc = jxlo.encode_vardct(img, **kw)
    out["enc" + str(sorted(kw.items()))] = hashlib.sha256(enc).hexdigest()
    out["dec" + str(sorted(kw.items()))] = hashlib.sha256(jxlo.decode(enc, 3, jxlo.UINT8).tobytes()).hexdigest()
    out["decf" + str(sorted(kw.items()))] = hashlib.sha256(jxlo.decode(enc, 3, jxlo.FLOAT).tobytes()).hexdigest()
print(json.dumps(out, sort_keys=True))
"""


def run(which):
    code = CHILD % (os.path.join(ROOT, "tests"), GOLDEN)
    r = subprocess.run([sys.executable, "-c", code, which], stdout=subprocess.PIPE, text=True, check=True)
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_fast_build_is_bit_identical():
    assert run("fast") == run("plain")
