"""Generates tests/golden/golden.json from the reference's own fixtures.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
Golden facts:
  G1  samples/sample.jxl == samples/sample.png as RGBA16 (asserted by jpegxl-rs/src/image.rs:158-174)
  G2  samples/bench.jxl  == samples/bench.png  as RGBA8  (the criterion bench input, jpegxl-rs/benches/decode.rs:10)
The PNGs are decoded here with OpenCV / PIL and only the SHA-256 of the raw pixel bytes is
committed for bench.png (2.4 MB); sample.png is small and committed as is.
"""
import hashlib
import json
import os

import cv2
import numpy as np
from PIL import Image

S = "/root/reference/samples/"
HERE = os.path.dirname(os.path.abspath(__file__))


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


g = {}
a = cv2.cvtColor(cv2.imread(S + "sample.png", cv2.IMREAD_UNCHANGED), cv2.COLOR_BGRA2RGBA)
assert a.dtype == np.uint16 and a.shape == (50, 40, 4)
g["sample.jxl"] = {"width": 40, "height": 50, "channels": 4, "dtype": "uint16", "sha256": sha(a),
                   "source": "samples/sample.png"}
b = np.array(Image.open(S + "bench.png").convert("RGBA"))
assert b.dtype == np.uint8 and b.shape == (1433, 2122, 4)
g["bench.jxl"] = {"width": 2122, "height": 1433, "channels": 4, "dtype": "uint8", "sha256": sha(b),
                  "source": "samples/bench.png",
                  "row_sha256_first": sha(b[0]), "row_sha256_last": sha(b[-1])}
for f in ["sample.jxl", "bench.jxl", "sample_grey.jxl", "sample_jpg.jxl", "2bit.jxl"]:
    g.setdefault(f, {})["file_sha256"] = hashlib.sha256(open(S + f, "rb").read()).hexdigest()
json.dump(g, open(os.path.join(HERE, "golden.json"), "w"), indent=1, sort_keys=True)
print(json.dumps(g, indent=1))
