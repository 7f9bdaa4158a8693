set -x
python bench.py --steps 8 --warmup 2 --inflight 1 --batch 64 --no-cpu-baseline > gpurun_out/bench_v12_b64_if1.json 2> gpurun_out/bench_v12.err; python tools/show_bench.py gpurun_out/bench_v12_b64_if1.json; tail -3 gpurun_out/bench_v12.err
python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v12_b256_if4.json 2> gpurun_out/bench_v12.err; python tools/show_bench.py gpurun_out/bench_v12_b256_if4.json; tail -3 gpurun_out/bench_v12.err
python bench.py --workload modular --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v12_mod.json 2> gpurun_out/bench_v12.err; python tools/show_bench.py gpurun_out/bench_v12_mod.json; tail -3 gpurun_out/bench_v12.err
# launch list of one vardct4k step (8 frames, one handle) and full captures of the two entropy kernels
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_v12_vardct4k_b8.csv python tools/ncu_workload.py 8 vardct_4k_natural.jxl 3 > gpurun_out/ncu_l.log 2>&1; tail -2 gpurun_out/ncu_l.log
for k in k_modular_decode_sparse k_ac_decode k_dequant_idct k_epf; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/r1_v12_$k \
    python tools/ncu_workload.py 8 vardct_4k_natural.jxl 3 > gpurun_out/ncu_$k.log 2>&1
  tail -2 gpurun_out/ncu_$k.log
done
