# warp-cooperative DC chain kernel (k_modular_decode_coop): parity, bench with / without, ncu of the kernel
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2i_pytest.log
run() { name=$1; shift; env "$@" python bench.py --steps 16 --warmup 2 --no-cpu-baseline --no-also > gpurun_out/r2i_$name.json 2> gpurun_out/r2i_$name.err; }
run coop
run nocoop JXLB200_NO_COOP=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_modular_decode_coop -s 1 -c 1 -f -o gpurun_out/r2i_ncu_coop \
  python tools/ncu_workload.py 8 vardct_4k_natural.jxl 3 > gpurun_out/r2i_ncu_coop.log 2>&1
tail -2 gpurun_out/r2i_ncu_coop.log
