set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v16_b256_if4.json 2> gpurun_out/bench_v16.err; python tools/show_bench.py gpurun_out/bench_v16_b256_if4.json; tail -4 gpurun_out/bench_v16.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_enc_emit -c 1 -f -o gpurun_out/r1_k_enc_emit_b8 python bench.py --workload encode4k --batch 8 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_emit.log 2>&1; tail -2 gpurun_out/ncu_emit.log
