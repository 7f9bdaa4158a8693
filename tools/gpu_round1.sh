set -x
python -m pytest tests/test_encoder.py tests/test_gpu_vardct.py -m gpu -x -q 2>&1 | tail -3
python bench.py --workload encode4k --batch 32 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_enc_v7_b32.json 2> gpurun_out/bench_enc.err; python tools/show_bench.py gpurun_out/bench_enc_v7_b32.json; tail -3 gpurun_out/bench_enc.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 16 -c 16 --csv --log-file gpurun_out/launches_enc_v7.csv python bench.py --workload encode4k --batch 8 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_enc.log 2>&1; tail -1 gpurun_out/ncu_enc.log | cut -c1-100
