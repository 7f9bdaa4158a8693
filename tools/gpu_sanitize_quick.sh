# memcheck of the decode kernels on two small lossy files (fits a one-minute gpurun call)
cat > /tmp/san.py <<'PY'
import sys, ctypes, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import __graft_entry__ as ge, vardct_cases as vc
pkg = ge.load_package()
files = [vc.encoded(n)[0] for n in ["heuristic", "odd_size", "all_strategies"]]
outs = pkg.decode_batch(files, 3, np.uint8)
print("ok", [o.shape for o in outs])
PY
timeout 55 compute-sanitizer --tool ${SAN_TOOL:-memcheck} --error-exitcode 3 python /tmp/san.py > gpurun_out/san_${SAN_TOOL:-memcheck}_v29.log 2>&1; echo ${SAN_TOOL:-memcheck} rc=$?; tail -3 gpurun_out/san_${SAN_TOOL:-memcheck}_v29.log
