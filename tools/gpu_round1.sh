set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r1_bench_final_default.json 2> gpurun_out/bench_final.err; python tools/show_bench.py gpurun_out/r1_bench_final_default.json | head -4; tail -1 gpurun_out/bench_final.err
