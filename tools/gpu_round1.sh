set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
nproc; free -g | head -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_b256.json 2> gpurun_out/bench_b256.err; tail -c 3000 gpurun_out/bench_b256.json
python bench.py --steps 3 --warmup 3 --batch 768 --no-cpu-baseline > gpurun_out/bench_b768.json 2> gpurun_out/bench_b768.err; tail -c 3000 gpurun_out/bench_b768.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python tools/ncu_workload.py 64 > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/ncu_launch.log
