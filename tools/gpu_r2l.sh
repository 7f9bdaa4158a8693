# per-pixel phases of concurrent handles take turns (PixelTurn): A/B with the lock-step phases, coop on / off, AC CTA widths
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2l_pytest.log
run() { name=$1; shift; env "$@" python bench.py --steps 24 --warmup 2 --no-cpu-baseline --no-also > gpurun_out/r2l_$name.json 2> gpurun_out/r2l_$name.err; }
run turns
run noturns JXLB200_PIXEL_TURNS=0
run turns_nocoop JXLB200_NO_COOP=1
run turns_acw4 JXLB200_AC_WARPS=4
run turns_acframe JXLB200_AC_FRAME=1
