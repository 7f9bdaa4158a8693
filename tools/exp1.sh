run() { name=$1; shift; env "$@" python bench.py --steps 16 --warmup 2 --no-cpu-baseline > gpurun_out/x1_$name.json 2> gpurun_out/x1_$name.err; }
run dcf64 JXLB200_DCF_THREADS=64
run dcf32 JXLB200_DCF_THREADS=32
run dcf64_w800 JXLB200_DCF_THREADS=64 JXLB200_WAVE_MB=800
run dcf64_w100 JXLB200_DCF_THREADS=64 JXLB200_WAVE_MB=100
run dcf64_w16g JXLB200_DCF_THREADS=64 JXLB200_WAVE_MB=16384
