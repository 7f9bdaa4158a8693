set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r1_bench_final_default.json 2> gpurun_out/bench_final.err; python tools/show_bench.py gpurun_out/r1_bench_final_default.json; tail -2 gpurun_out/bench_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1_bench_final_reference.json 2>> gpurun_out/bench_final.err; cut -c1-200 gpurun_out/r1_bench_final_reference.json
python bench.py --workload encode4k --batch 32 --steps 3 --warmup 1 > gpurun_out/r1_bench_final_encode.json 2>> gpurun_out/bench_final.err; python tools/show_bench.py gpurun_out/r1_bench_final_encode.json
python bench.py --workload modular --steps 8 --warmup 2 > gpurun_out/r1_bench_final_modular.json 2>> gpurun_out/bench_final.err; python tools/show_bench.py gpurun_out/r1_bench_final_modular.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1_launches_bench_default_b256_final.csv python bench.py --steps 1 --warmup 1 --inflight 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
