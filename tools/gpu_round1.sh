set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
nproc; free -g | head -2
python bench.py --steps 5 --warmup 3 --cpu-seconds 20 > gpurun_out/bench_vardct4k_b64.json 2> gpurun_out/bench_vardct4k_b64.err; tail -c 4000 gpurun_out/bench_vardct4k_b64.json; tail -5 gpurun_out/bench_vardct4k_b64.err
python bench.py --workload modular --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_modular_b256.json 2> gpurun_out/bench_modular_b256.err; tail -c 3000 gpurun_out/bench_modular_b256.json; tail -5 gpurun_out/bench_modular_b256.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_vardct4k.csv python bench.py --steps 1 --warmup 1 --batch 16 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/ncu_launch.log
