set -x
# one full capture per kernel of the VarDCT path (8 frames, second Run = warm), replayed by ncu ~40x each
for k in k_modular_decode k_ac_decode k_dequant_idct k_epf k_color_write k_dc_finish k_gaborish; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/r1_$k \
    python tools/ncu_workload.py 8 vardct_4k_natural.jxl 3 > gpurun_out/ncu_$k.log 2>&1
  tail -2 gpurun_out/ncu_$k.log
done
ls -la gpurun_out
