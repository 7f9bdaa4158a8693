set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v24_if4.json 2> gpurun_out/bench_v24.err; python tools/show_bench.py gpurun_out/bench_v24_if4.json; tail -1 gpurun_out/bench_v24.err
python bench.py --steps 16 --warmup 3 --inflight 8 --no-cpu-baseline > gpurun_out/bench_v24_if8.json 2> gpurun_out/bench_v24.err; python tools/show_bench.py gpurun_out/bench_v24_if8.json; tail -1 gpurun_out/bench_v24.err
JXLB200_NO_PRIORITY_STREAM=1 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v24_if4_noprio.json 2> gpurun_out/bench_v24.err; python tools/show_bench.py gpurun_out/bench_v24_if4_noprio.json; tail -1 gpurun_out/bench_v24.err
