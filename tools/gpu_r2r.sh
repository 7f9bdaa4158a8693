# splines (2bit.jxl) on the GPU + everything else; traffic of one decode step for roofline.traffic
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2r_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2r_smoke.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv \
  --log-file gpurun_out/r2r_launches_traffic.csv python bench.py --steps 1 --warmup 1 --inflight 1 --no-cpu-baseline --no-also > gpurun_out/r2r_ncu_bench.log 2>&1
