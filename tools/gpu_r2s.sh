# k_render_fused: last filter stage fused with the colour / store loop (4 pixels per thread): parity, bench
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2s_pytest.log
python bench.py --steps 16 --warmup 2 --no-cpu-baseline --no-also > gpurun_out/r2s_fused_last.json 2> gpurun_out/r2s_fused_last.err
