# racecheck again after the per-sample __syncwarp in the coop kernel, its cost alone (16 different 4K frames, one handle)
sed -i 's/for tool in memcheck racecheck/for tool in racecheck/' tools/gpu_sanitize_r2.sh
bash tools/gpu_sanitize_r2.sh > gpurun_out/r2p_san.log 2>&1
python tools/exp_distinct.py 16 > gpurun_out/r2p_distinct16.jsonl 2> gpurun_out/r2p_distinct16.err
JXLB200_NO_COOP=1 python tools/exp_distinct.py 16 > gpurun_out/r2p_distinct16_nocoop.jsonl 2>> gpurun_out/r2p_distinct16.err
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2p_pytest.log
