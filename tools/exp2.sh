run() { name=$1; shift; env "$@" python tools/interference.py --quick --bg-handles 3 > gpurun_out/x3_intf_$name.jsonl 2> gpurun_out/x3_intf_$name.err; env "$@" python bench.py --steps 16 --warmup 2 --no-cpu-baseline > gpurun_out/x3_$name.json 2> gpurun_out/x3_$name.err; }
run a8_d64 JXLB200_AC_WARPS=8 JXLB200_DCF_THREADS=64
run a8_d64_dense JXLB200_AC_WARPS=8 JXLB200_DCF_THREADS=64 JXLB200_MODULAR_DENSE=1
run a4_d32 JXLB200_AC_WARPS=4 JXLB200_DCF_THREADS=32
