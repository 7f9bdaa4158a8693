set -x
python bench.py > gpurun_out/bench_final_default.json 2> gpurun_out/bench_final.err; python tools/show_bench.py gpurun_out/bench_final_default.json; tail -3 gpurun_out/bench_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_reference.json 2> gpurun_out/bench_final.err; cut -c1-400 gpurun_out/bench_final_reference.json; tail -3 gpurun_out/bench_final.err
python bench.py --workload modular > gpurun_out/bench_final_modular.json 2> gpurun_out/bench_final.err; python tools/show_bench.py gpurun_out/bench_final_modular.json; tail -3 gpurun_out/bench_final.err
python bench.py --workload encode4k --batch 32 --steps 4 --warmup 1 > gpurun_out/bench_final_encode.json 2> gpurun_out/bench_final.err; python tools/show_bench.py gpurun_out/bench_final_encode.json; tail -3 gpurun_out/bench_final.err
python -c "import __graft_entry__ as g; g.smoke()"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_modular_decode_sparse -s 1 -c 1 -f -o gpurun_out/r1_final_k_modular_decode_sparse_b256 python tools/ncu_workload.py 256 vardct_4k_natural.jxl 3 > gpurun_out/ncu_final.log 2>&1; tail -2 gpurun_out/ncu_final.log
