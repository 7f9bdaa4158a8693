# lossy encode with alpha + everything else: parity, encode bench (no regression), default bench
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2o_pytest.log
python bench.py --workload encode4k --batch 32 --steps 3 --warmup 1 --no-cpu-baseline --no-also > gpurun_out/r2o_encode.json 2> gpurun_out/r2o_encode.err
