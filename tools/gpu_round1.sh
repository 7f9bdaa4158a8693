set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --inflight 1 > gpurun_out/bench_v3_b64_if1.json 2> gpurun_out/bench_v3.err; python tools/show_bench.py gpurun_out/bench_v3_b64_if1.json; tail -3 gpurun_out/bench_v3.err
python bench.py --steps 6 --warmup 2 --no-cpu-baseline --inflight 2 --batch 256 > gpurun_out/bench_v3_b256_if2.json 2> gpurun_out/bench_v3.err; python tools/show_bench.py gpurun_out/bench_v3_b256_if2.json; tail -3 gpurun_out/bench_v3.err
python bench.py --workload modular --steps 3 --warmup 2 --no-cpu-baseline --inflight 1 > gpurun_out/bench_v3_mod.json 2> gpurun_out/bench_v3.err; python tools/show_bench.py gpurun_out/bench_v3_mod.json; tail -3 gpurun_out/bench_v3.err
