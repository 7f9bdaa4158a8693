set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 6 --warmup 3 --inflight 1 --no-cpu-baseline > gpurun_out/bench_v6_b64_if1.json 2> gpurun_out/bench_v6.err; python tools/show_bench.py gpurun_out/bench_v6_b64_if1.json; tail -3 gpurun_out/bench_v6.err
python bench.py --steps 8 --warmup 3 --inflight 2 --no-cpu-baseline > gpurun_out/bench_v6_b64_if2.json 2> gpurun_out/bench_v6.err; python tools/show_bench.py gpurun_out/bench_v6_b64_if2.json; tail -3 gpurun_out/bench_v6.err
python bench.py --steps 9 --warmup 3 --inflight 3 --no-cpu-baseline > gpurun_out/bench_v6_b64_if3.json 2> gpurun_out/bench_v6.err; python tools/show_bench.py gpurun_out/bench_v6_b64_if3.json; tail -3 gpurun_out/bench_v6.err
python bench.py --steps 6 --warmup 2 --inflight 2 --batch 256 --no-cpu-baseline > gpurun_out/bench_v6_b256_if2.json 2> gpurun_out/bench_v6.err; python tools/show_bench.py gpurun_out/bench_v6_b256_if2.json; tail -3 gpurun_out/bench_v6.err
