# ncu captures used for profiles/: full sets of the two per-pixel kernels on 8 x vardct_4k_natural.jxl, and the launch
# list of the bench command (one handle, one step). Summaries: tools/ncu_summary.py, tools/ncu_lines.py, tools/launch_shares.py
set -x
for k in k_render_fused k_dequant_idct; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/ncu_$k \
    python tools/ncu_workload.py 8 vardct_4k_natural.jxl 3 > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_default.csv \
  python bench.py --steps 1 --warmup 1 --inflight 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-200
