# Round 2, call d: parity + memcheck of the new kernels, then A/B: CTA-per-frame AC decode vs the one-warp kernel.
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2d_pytest.log
cat > /tmp/san.py <<'PY'
import sys, ctypes, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import __graft_entry__ as ge, vardct_cases as vc
pkg = ge.load_package()
files = [vc.encoded(n)[0] for n in ["heuristic", "odd_size", "all_strategies", "three_passes"]]
outs = pkg.decode_batch(files, 3, np.uint8)
print("ok", [o.shape for o in outs])
PY
timeout 120 compute-sanitizer --tool memcheck --error-exitcode 3 python /tmp/san.py > gpurun_out/r2d_memcheck.log 2>&1; echo memcheck rc=$? >> gpurun_out/r2d_memcheck.log
run() { name=$1; shift; env "$@" python bench.py --steps 16 --warmup 2 --no-cpu-baseline --no-also > gpurun_out/r2d_$name.json 2> gpurun_out/r2d_$name.err; }
run base
run onewarp JXLB200_AC_ONE_WARP=1
python tools/interference.py --quick --bg-handles 3 > gpurun_out/r2d_intf.jsonl 2> gpurun_out/r2d_intf.err
