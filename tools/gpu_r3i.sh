# the two emission kernels after the split: launch times and a full capture of the DC-half kernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r3i_launches.csv -k regex:'k_enc_emit' python tools/ncu_workload_enc.py 16 > gpurun_out/r3i_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_enc_emit_dc' -c 1 -o gpurun_out/r3i_emit_dc -f python tools/ncu_workload_enc.py 16 > gpurun_out/r3i_ncu2.log 2>&1
grep -h "k_enc_emit" gpurun_out/r3i_launches.csv | cut -d, -f5,12- | head
