set -x
python bench.py --workload modular > gpurun_out/r1_bench_final_modular.json 2> gpurun_out/bench_final_mod.err; python tools/show_bench.py gpurun_out/r1_bench_final_modular.json; tail -3 gpurun_out/bench_final_mod.err
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v26_s4.json 2> gpurun_out/bench_v26.err; python tools/show_bench.py gpurun_out/bench_v26_s4.json | head -1; tail -1 gpurun_out/bench_v26.err
