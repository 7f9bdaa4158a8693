set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v19_b256_if4.json 2> gpurun_out/bench_v19.err; python tools/show_bench.py gpurun_out/bench_v19_b256_if4.json; tail -2 gpurun_out/bench_v19.err
JXLB200_UNFUSED_RENDER=1 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v19_b256_if4_unfused.json 2> gpurun_out/bench_v19.err; python tools/show_bench.py gpurun_out/bench_v19_b256_if4_unfused.json; tail -2 gpurun_out/bench_v19.err
python bench.py --steps 4 --warmup 2 --inflight 1 --batch 64 --no-cpu-baseline > gpurun_out/bench_v19_b64_if1.json 2> gpurun_out/bench_v19.err; python tools/show_bench.py gpurun_out/bench_v19_b64_if1.json; tail -2 gpurun_out/bench_v19.err
