python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2g_pytest.log
python tools/exp_distinct.py 128 > gpurun_out/r2g_distinct.jsonl 2> gpurun_out/r2g_distinct.err
JXLB200_AC_FRAME=1 python tools/exp_distinct.py 128 > gpurun_out/r2g_distinct_frame.jsonl 2> gpurun_out/r2g_distinct_frame.err
run() { name=$1; shift; env "$@" python bench.py --steps 16 --warmup 2 --no-cpu-baseline --no-also > gpurun_out/r2g_$name.json 2> gpurun_out/r2g_$name.err; }
run base
run frame JXLB200_AC_FRAME=1
