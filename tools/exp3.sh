python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/x4_pytest.log
run() { name=$1; shift; env "$@" python bench.py --steps 16 --warmup 2 --no-cpu-baseline > gpurun_out/x4_$name.json 2> gpurun_out/x4_$name.err; }
run base
run a8 JXLB200_AC_WARPS=8
python tools/interference.py --quick --bg-handles 3 > gpurun_out/x4_intf.jsonl 2> gpurun_out/x4_intf.err
