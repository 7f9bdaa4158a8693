# compute-sanitizer memcheck + racecheck of the kernels and calls added at the end of round 2: the warp-per-section rANS
# writers (atomics into shared words, shuffles), the split DC-group sections, k_enc_compact, the per-tile strategy
# kernel, the lossless encoder at 128 x 128 groups, splines on lossy frames, and the streaming calls
# (PlanBatch / CommitPlan / RunToHost with the per-wave copies on the copy stream)
cat > /tmp/san2.py <<'PY'
import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import torch
import __graft_entry__ as ge, vardct_cases as vc, jxlo
pkg = ge.load_package()
img = vc.crop(300, 520, 100, 200)
a = ((img[:, :, 0].astype(np.int32) + np.arange(520)[None, :] * 3) % 256).astype(np.uint8)
rgba = np.dstack([img, a])
l = pkg.JxlEncoder(lossless=True, uses_original_profile=True, has_alpha=True).encode_batch([rgba, rgba[:100, :128]])
e = pkg.JxlEncoder(has_alpha=True).encode_batch([rgba, np.ascontiguousarray(rgba[:200, :256])])
e3 = pkg.JxlEncoder().encode_batch([img])
assert l[0].data == jxlo.encode_modular(rgba.astype(np.uint16), tree=1, predictor=5, group_size_shift=0, bits=8, alpha=True, rct=6)
files = [l[0].data, l[1].data, e[0].data, e[1].data, e3[0].data, jxlo.encode_vardct(img, strategy_mode=2, splines=5),
         jxlo.encode_vardct(rgba[:60, :70], strategy_mode=2, splines=3)]
want = [jxlo.decode(f, 4, jxlo.UINT8) for f in files]
assert np.array_equal(want[0], rgba)
d = pkg.BatchDecoder(0)
stream = torch.cuda.Stream()
d.plan(files, 4, pkg.JXL_TYPE_UINT8)
for step in range(3):
    d.commit()
    outs = [torch.empty(d.out_size(i), dtype=torch.uint8).pin_memory().numpy() for i in range(len(files))]
    d.run_to_host(outs, stream.cuda_stream)
    d.plan(files, 4, pkg.JXL_TYPE_UINT8)
    d.wait(stream.cuda_stream)
    for o, w in zip(outs, want):
        assert np.array_equal(o.reshape(w.shape), w)
print("ok", len(files), "files, encoders byte-exact, streaming decode bit-exact")
PY
for tool in memcheck racecheck; do
  timeout 400 compute-sanitizer --tool $tool --error-exitcode 3 python /tmp/san2.py > gpurun_out/r2b_san_$tool.log 2>&1; echo $tool rc=$?; tail -4 gpurun_out/r2b_san_$tool.log
done
