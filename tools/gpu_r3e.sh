# encoder launch list (ncu, 16 4K frames encoded once by tools/ncu_workload_enc.py) + one full capture of the warp emit kernel;
# then the two-GPU record runs (our arm and the reference arm)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/r3e_launches_enc.csv -k regex:'k_enc|k_encl' python tools/ncu_workload_enc.py 16 > gpurun_out/r3e_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_enc_emit' -c 1 -o gpurun_out/r3e_emit -f \
  python tools/ncu_workload_enc.py 16 > gpurun_out/r3e_ncu2.log 2>&1
tail -2 gpurun_out/r3e_ncu2.log
