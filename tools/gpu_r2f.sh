python tools/exp_distinct.py 128 > gpurun_out/r2f_distinct.jsonl 2> gpurun_out/r2f_distinct.err
