set -x
JXLB200_MODULAR_DENSE=1 python bench.py --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v27_shared_dense.json 2> gpurun_out/bench_v27.err; python tools/show_bench.py gpurun_out/bench_v27_shared_dense.json | head -2; tail -1 gpurun_out/bench_v27.err
