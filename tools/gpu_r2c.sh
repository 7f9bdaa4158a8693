# ncu --set full captures of the three kernels that carry the step's instruction volume, on 8 x vardct_4k_natural.jxl
for k in k_render_fused k_dequant_idct k_ac_decode; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/r2c_ncu_$k \
    python tools/ncu_workload.py 8 vardct_4k_natural.jxl 3 > gpurun_out/r2c_ncu_$k.log 2>&1
  tail -1 gpurun_out/r2c_ncu_$k.log
done
