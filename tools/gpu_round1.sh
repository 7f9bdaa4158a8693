set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v21_if4.json 2> gpurun_out/bench_v21.err; python tools/show_bench.py gpurun_out/bench_v21_if4.json; tail -1 gpurun_out/bench_v21.err
python bench.py --steps 16 --warmup 3 --inflight 8 --no-cpu-baseline > gpurun_out/bench_v21_if8.json 2> gpurun_out/bench_v21.err; python tools/show_bench.py gpurun_out/bench_v21_if8.json; tail -1 gpurun_out/bench_v21.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/bench_v21_if4.json").read().strip().splitlines()[-1])
print("alone:", {k:round(v,1) for k,v in j["roofline"]["all_kernels_ms_one_handle_alone"].items() if v>0.01})
PY
