set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v18_b256_if4.json 2> gpurun_out/bench_v18.err; python tools/show_bench.py gpurun_out/bench_v18_b256_if4.json; tail -4 gpurun_out/bench_v18.err
python bench.py --steps 12 --warmup 3 --inflight 8 --no-cpu-baseline > gpurun_out/bench_v18_b256_if8.json 2> gpurun_out/bench_v18.err; python tools/show_bench.py gpurun_out/bench_v18_b256_if8.json; tail -4 gpurun_out/bench_v18.err
python bench.py --steps 6 --warmup 3 --inflight 3 --batch 768 --no-cpu-baseline > gpurun_out/bench_v18_b768_if3.json 2> gpurun_out/bench_v18.err; python tools/show_bench.py gpurun_out/bench_v18_b768_if3.json; tail -4 gpurun_out/bench_v18.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
nproc; free -g | head -2
