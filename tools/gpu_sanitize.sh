set -x
cat > /tmp/san.py <<'PY'
import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import __graft_entry__ as ge, vardct_cases as vc
from conftest import read_golden
pkg = ge.load_package()
names = ["heuristic", "all_strategies", "three_passes", "odd_size", "strategy_18", "strategy_24"]
files = [vc.encoded(n)[0] for n in names] + [read_golden("sample.jxl")]
outs = pkg.decode_batch(files, 3, np.uint8)
enc = pkg.encoder_builder().build()
res = enc.encode_batch([vc.crop(300, 400), vc.crop(200, 200)])
print("ok", [o.shape for o in outs], [len(r.data) for r in res])
PY
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 python /tmp/san.py > gpurun_out/san_memcheck.log 2>&1; echo memcheck rc=$?; tail -4 gpurun_out/san_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 3 python /tmp/san.py > gpurun_out/san_racecheck.log 2>&1; echo racecheck rc=$?; tail -6 gpurun_out/san_racecheck.log
