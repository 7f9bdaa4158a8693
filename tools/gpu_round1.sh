set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v17_b256_if4.json 2> gpurun_out/bench_v17.err; python tools/show_bench.py gpurun_out/bench_v17_b256_if4.json; tail -4 gpurun_out/bench_v17.err
python bench.py --steps 8 --warmup 2 --inflight 1 --batch 64 --no-cpu-baseline > gpurun_out/bench_v17_b64_if1.json 2> gpurun_out/bench_v17.err; python tools/show_bench.py gpurun_out/bench_v17_b64_if1.json; tail -4 gpurun_out/bench_v17.err
python bench.py --workload encode4k --batch 32 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_enc_v8_b32.json 2> gpurun_out/bench_enc.err; python tools/show_bench.py gpurun_out/bench_enc_v8_b32.json; tail -3 gpurun_out/bench_enc.err
