set -x
python -m pytest tests/test_encoder.py -m gpu -x -q 2>&1 | tail -3
python bench.py --workload encode4k --batch 32 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_enc_v11_b32.json 2> gpurun_out/bench_enc.err; python tools/show_bench.py gpurun_out/bench_enc_v11_b32.json; tail -3 gpurun_out/bench_enc.err
