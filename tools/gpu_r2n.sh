# alpha extra channel in lossy frames (chained Modular streams, second launch) + everything else: parity, default bench
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2n_pytest.log
python bench.py --steps 16 --warmup 2 --no-cpu-baseline --no-also > gpurun_out/r2n_default.json 2> gpurun_out/r2n_default.err
