# Round 2, first GPU call: parity tests, smoke, the default bench line (with `also`), the reference arm, and the
# launch list + DRAM traffic of all kernels of one step (one handle).
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2a_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1
python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2a_ref.json 2> gpurun_out/r2a_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 800 --csv \
  --log-file gpurun_out/r2a_launches_traffic.csv python bench.py --steps 1 --warmup 1 --inflight 1 --no-cpu-baseline --no-also \
  > gpurun_out/r2a_ncu_bench.log 2>&1
nvidia-smi topo -m > gpurun_out/r2a_topo.txt 2>&1; nproc >> gpurun_out/r2a_topo.txt; free -g >> gpurun_out/r2a_topo.txt
