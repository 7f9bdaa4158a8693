set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 8 --warmup 2 --inflight 1 --batch 64 --no-cpu-baseline > gpurun_out/bench_v11_b64_if1.json 2> gpurun_out/bench_v11.err; python tools/show_bench.py gpurun_out/bench_v11_b64_if1.json; tail -3 gpurun_out/bench_v11.err
