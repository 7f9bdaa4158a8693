set -x
python -m pytest tests/test_gpu_decode.py tests/test_gpu_vardct.py -m gpu -x -q 2>&1 | tail -3
for cfg in "64 1" "256 2" "256 4"; do
set -- $cfg
python bench.py --steps 8 --warmup 2 --inflight $2 --batch $1 --no-cpu-baseline > gpurun_out/bench_v10_b$1_if$2.json 2> gpurun_out/bench_v10.err; python tools/show_bench.py gpurun_out/bench_v10_b$1_if$2.json; tail -3 gpurun_out/bench_v10.err
done
python bench.py --workload modular --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/bench_v10_mod.json 2> gpurun_out/bench_v10.err; python tools/show_bench.py gpurun_out/bench_v10_mod.json; tail -3 gpurun_out/bench_v10.err
