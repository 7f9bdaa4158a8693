# warp-cooperative DC chain kernel, second version (fast rows, 4 warps per CTA): parity, bench with / without, ncu
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2j_pytest.log
run() { name=$1; shift; env "$@" python bench.py --steps 16 --warmup 2 --no-cpu-baseline --no-also > gpurun_out/r2j_$name.json 2> gpurun_out/r2j_$name.err; }
run coop
run nocoop JXLB200_NO_COOP=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_modular_decode_coop -s 1 -c 1 -f -o gpurun_out/r2j_ncu_coop \
  python tools/ncu_workload_enc.py 16 > gpurun_out/r2j_ncu_coop.log 2>&1
tail -2 gpurun_out/r2j_ncu_coop.log
