set -x
python bench.py --steps 6 --warmup 2 --inflight 2 --batch 256 --no-cpu-baseline > gpurun_out/bench_v8_b256_if2.json 2> gpurun_out/bench_v8.err; python tools/show_bench.py gpurun_out/bench_v8_b256_if2.json; tail -3 gpurun_out/bench_v8.err
python bench.py --steps 6 --warmup 2 --inflight 3 --batch 128 --no-cpu-baseline > gpurun_out/bench_v8_b128_if3.json 2> gpurun_out/bench_v8.err; python tools/show_bench.py gpurun_out/bench_v8_b128_if3.json; tail -3 gpurun_out/bench_v8.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_v8_ref.json 2> gpurun_out/bench_v8.err; cat gpurun_out/bench_v8_ref.json | cut -c1-600; tail -3 gpurun_out/bench_v8.err
python -c "import __graft_entry__ as g; g.smoke()"
