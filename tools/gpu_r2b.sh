# Round 2, call b: parity, then A/B runs of the carve-out preference, the (y, N, W) table path and the special-transform path.
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2b_pytest.log
run() { name=$1; shift; env "$@" python bench.py --steps 16 --warmup 2 --no-cpu-baseline --no-also > gpurun_out/r2b_$name.json 2> gpurun_out/r2b_$name.err; }
run base
run nocarve JXLB200_NO_CARVEOUT=1
run nonw JXLB200_NO_NW_LUT=1
python tools/interference.py --quick --bg-handles 3 > gpurun_out/r2b_intf.jsonl 2> gpurun_out/r2b_intf.err
