# the round's record runs: smoke, default bench (CPU baseline + side runs), reference arm, ncu launch list of the bench command, racecheck
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2q_smoke.log 2>&1
python bench.py > gpurun_out/r2q_default.json 2> gpurun_out/r2q_default.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2q_reference.json 2> gpurun_out/r2q_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2q_launches.csv \
  python bench.py --steps 1 --warmup 1 --inflight 1 --no-cpu-baseline --no-also > gpurun_out/r2q_ncu_bench.log 2>&1
sed -i 's/for tool in memcheck racecheck/for tool in racecheck/' tools/gpu_sanitize_r2.sh
bash tools/gpu_sanitize_r2.sh > gpurun_out/r2q_san.log 2>&1
