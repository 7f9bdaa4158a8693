# warp-cooperative DC chain kernel, third version (fast rows also with clusters outside the lanes): parity, bench
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2k_pytest.log
run() { name=$1; shift; env "$@" python bench.py --steps 16 --warmup 2 --no-cpu-baseline --no-also > gpurun_out/r2k_$name.json 2> gpurun_out/r2k_$name.err; }
run coop
python bench.py --steps 18 --warmup 2 --inflight 6 --no-cpu-baseline --no-also > gpurun_out/r2k_coop_if6.json 2> gpurun_out/r2k_coop_if6.err
python bench.py --steps 16 --warmup 2 --inflight 4 --no-cpu-baseline --no-also > gpurun_out/r2k_coop_if4.json 2> gpurun_out/r2k_coop_if4.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_modular_decode_coop -s 1 -c 1 -f -o gpurun_out/r2k_ncu_coop \
  python tools/ncu_workload_enc.py 16 > gpurun_out/r2k_ncu_coop.log 2>&1
tail -2 gpurun_out/r2k_ncu_coop.log
