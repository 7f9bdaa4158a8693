set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --no-cpu-baseline > gpurun_out/bench_v29_wave.json 2> gpurun_out/bench_v29.err; python tools/show_bench.py gpurun_out/bench_v29_wave.json | head -2; tail -1 gpurun_out/bench_v29.err
