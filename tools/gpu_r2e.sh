# Round 2, call e: SM partition (green contexts) A/B.
run() { name=$1; shift; env "$@" python bench.py --steps 16 --warmup 2 --no-cpu-baseline --no-also > gpurun_out/r2e_$name.json 2> gpurun_out/r2e_$name.err; }
run base
run sm24 JXLB200_ENTROPY_SMS=24
run sm32 JXLB200_ENTROPY_SMS=32
run sm48 JXLB200_ENTROPY_SMS=48
JXLB200_ENTROPY_SMS=32 python -m pytest tests/test_gpu_vardct.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2e_pytest_sm32.log
