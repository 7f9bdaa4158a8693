# One gpurun call that checks the tree on a B200: parity tests, smoke, the default bench line.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_round1.sh'
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; python tools/show_bench.py gpurun_out/bench_default.json | head -4; tail -2 gpurun_out/bench_default.err
