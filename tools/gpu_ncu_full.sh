set -x
for k in k_modular_decode_sparse k_ac_decode; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/r1_v8_$k \
    python tools/ncu_workload.py 8 vardct_4k_natural.jxl 3 > gpurun_out/ncu_$k.log 2>&1
  tail -2 gpurun_out/ncu_$k.log
done
