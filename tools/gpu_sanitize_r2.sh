# compute-sanitizer memcheck + racecheck of the round-2 kernels on small inputs: the warp-cooperative DC chain kernel
# (fast rows, generic rows with clusters outside the lanes), the chained Modular streams of an alpha frame, upsampling,
# orientation stores, the lossless encoder, the lossy encoder with alpha; and memcheck of a truncated / bit-flipped file
# (the device bit reader must stay inside the byte pool: ADVICE r1, high)
cat > /tmp/san.py <<'PY'
import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import __graft_entry__ as ge, vardct_cases as vc, jxlo
pkg = ge.load_package()
img = vc.crop(300, 520, 100, 200)
a = ((img[:, :, 0].astype(np.int32) + np.arange(520)[None, :] * 3) % 256).astype(np.uint8)
rgba = np.dstack([img, a])
files = [vc.encoded(n)[0] for n in ["heuristic", "odd_size", "all_strategies", "upsampling_2"]]
files += [jxlo.encode_vardct(img, strategy_mode=2, dc_tree=1), jxlo.encode_vardct(rgba, strategy_mode=2),
          jxlo.encode_vardct(rgba[:200, :256], strategy_mode=2, orientation=6)]
outs = pkg.decode_batch(files, 4, np.uint8)
enc = pkg.JxlEncoder(lossless=True, uses_original_profile=True, has_alpha=True)
l = enc.encode_batch([rgba])[0].data
e = pkg.JxlEncoder(has_alpha=True).encode_batch([rgba])[0].data
back = pkg.decode_batch([l, e], 4, np.uint8)
assert np.array_equal(back[0], rgba)
bad = 0
for cut in (0.5, 0.9):
    for f in (files[4], files[5]):
        b = bytearray(f[:int(len(f) * cut)] if cut < 0.9 else f)
        if cut >= 0.9:
            for k in range(len(b) // 2, len(b), 97): b[k] ^= 0x55
        try:
            pkg.decode_batch([bytes(b)], 4, np.uint8)
        except Exception:
            bad += 1
print("ok", [o.shape for o in outs], "corrupt inputs refused:", bad, "of 4")
PY
for tool in memcheck racecheck; do
  timeout 280 compute-sanitizer --tool $tool --error-exitcode 3 python /tmp/san.py > gpurun_out/r2_san_$tool.log 2>&1; echo $tool rc=$?; tail -4 gpurun_out/r2_san_$tool.log
done
