set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v23_if4.json 2> gpurun_out/bench_v23.err; python tools/show_bench.py gpurun_out/bench_v23_if4.json; tail -1 gpurun_out/bench_v23.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/bench_v23_if4.json").read().strip().splitlines()[-1])
print("alone:", {k:round(v,1) for k,v in j["roofline"]["all_kernels_ms_one_handle_alone"].items() if v>0.01})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_vardct4k_b8_v23.csv python tools/ncu_workload.py 8 vardct_4k_natural.jxl 3 > gpurun_out/ncu_launches.log 2>&1
grep -E "k_idct_mid|k_dequant_idct|k_render_fused" gpurun_out/r1_launches_vardct4k_b8_v23.csv | head -4 | cut -d, -f5,15
