set -x
for k in k_render_fused k_dequant_idct; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/r1_v19_$k \
    python tools/ncu_workload.py 8 vardct_4k_natural.jxl 3 > gpurun_out/ncu_$k.log 2>&1
  tail -2 gpurun_out/ncu_$k.log
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_vardct4k_b8_v19.csv python tools/ncu_workload.py 8 vardct_4k_natural.jxl 3 > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
