set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v28_bitmap.json 2> gpurun_out/bench_v28.err; python tools/show_bench.py gpurun_out/bench_v28_bitmap.json | head -2; tail -1 gpurun_out/bench_v28.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/bench_v28_bitmap.json").read().strip().splitlines()[-1])
print("alone:", {k:round(v,1) for k,v in j["roofline"]["all_kernels_ms_one_handle_alone"].items() if v>0.01})
PY
