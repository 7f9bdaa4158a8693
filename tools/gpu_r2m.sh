# orientation + lossless encoder + final coop policy: parity, default bench, lossless encode throughput, ncu of the coop kernel
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2m_pytest.log
python tools/bench_lossless_enc.py 8 3 > gpurun_out/r2m_lossless_enc.json 2> gpurun_out/r2m_lossless_enc.err
python bench.py --steps 16 --warmup 2 --no-cpu-baseline --no-also > gpurun_out/r2m_default.json 2> gpurun_out/r2m_default.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_modular_decode_coop -s 1 -c 1 -f -o gpurun_out/r2m_ncu_coop \
  python tools/ncu_workload_enc.py 16 > gpurun_out/r2m_ncu_coop.log 2>&1
tail -2 gpurun_out/r2m_ncu_coop.log
