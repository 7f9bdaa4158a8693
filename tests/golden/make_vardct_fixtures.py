"""Writes the 4K VarDCT fixtures of the bench / parity tests with the oracle's plain encoder and records the
sha256 of their oracle-decoded RGB8 pixels in golden.json. The reference ships no VarDCT file larger than 40x50
and libjxl cannot be built in this image, so these streams are ours; their content is the
reference's bench image tiled to 3840x2160 (natural statistics) and a seeded procedural frame.

    python tests/golden/make_vardct_fixtures.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import jxlo  # noqa: E402
import vardct_cases as vc  # noqa: E402


def main():
    gpath = os.path.join(HERE, "golden.json")
    g = json.load(open(gpath))
    # What a libjxl effort-7 stream at distance 1.0 exercises: block sizes and the 8x8 special transforms chosen by
    # libjxl's entropy-estimate search (strategy_mode 3, oracle/jxlo_enc_acs.h), adaptive quantisation field, fitted
    # chroma-from-luma maps, custom coefficient orders, inverse Gaborish, Gaborish + one EPF iteration in the loop
    # filter (lib/jxl/enc_frame.cc:254-287: epf_iters = 1 for 0.7 <= distance < 1.5), the fixed weighted-predictor DC tree.
    kw_e7 = dict(distance=1.0, strategy_mode=3, epf_iters=1)
    frames = {"vardct_4k_natural.jxl": (vc.frame_4k(), kw_e7),
              "vardct_4k_synthetic.jxl": (vc.synthetic(2160, 3840, 0xB200), kw_e7)}
    for name, (img, kw) in frames.items():
        data = jxlo.encode_vardct(img, **kw)
        open(os.path.join(HERE, name), "wb").write(data)
        px = jxlo.decode(data, 3, jxlo.UINT8)
        err = px.astype(float) - img
        g[name] = {"file_sha256": hashlib.sha256(data).hexdigest(), "sha256_rgb8": hashlib.sha256(px.tobytes()).hexdigest(), "width": 3840, "height": 2160,
                   "bytes": len(data), "bpp": round(8 * len(data) / (3840 * 2160), 4),
                   "psnr_vs_source": round(float(10 * __import__("numpy").log10(255 ** 2 / (err ** 2).mean())), 2),
                   "encoder": "oracle/jxlo_encode.h " + json.dumps(kw, sort_keys=True)}
        # what the encoder bench feeds back in: the decoded frame, encoded again (gradient DC tree, as the GPU encoder)
        again = jxlo.encode_vardct(px, distance=1.0, strategy_mode=2, dc_tree=1)
        g[name]["reencoded_sha256"] = hashlib.sha256(again).hexdigest()
        g[name]["reencoded_bytes"] = len(again)
        print(name, g[name])
    json.dump(g, open(gpath, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
