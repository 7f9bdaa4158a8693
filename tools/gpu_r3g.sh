# full captures of the encoder's forward-transform and chroma-from-luma kernels (16 4K frames)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_enc_coeffs' -c 1 -o gpurun_out/r3g_coeffs0 -f \
  python tools/ncu_workload_enc.py 16 > gpurun_out/r3g_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_enc_cfl' -c 1 -o gpurun_out/r3g_cfl -f \
  python tools/ncu_workload_enc.py 16 > gpurun_out/r3g_ncu2.log 2>&1
tail -1 gpurun_out/r3g_ncu2.log
