set -x
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v15_b256_if4.json 2> gpurun_out/bench_v15.err; python tools/show_bench.py gpurun_out/bench_v15_b256_if4.json; tail -4 gpurun_out/bench_v15.err
python bench.py --steps 8 --warmup 3 --inflight 2 --no-cpu-baseline > gpurun_out/bench_v15_b256_if2.json 2> gpurun_out/bench_v15.err; python tools/show_bench.py gpurun_out/bench_v15_b256_if2.json; tail -4 gpurun_out/bench_v15.err
nproc; free -g | head -2
