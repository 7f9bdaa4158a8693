JXLB200_RUN_TIMING=1 JXLB200_SPIN_WAIT=1 python bench.py --no-cpu-baseline --no-also --steps 8 > gpurun_out/r3a.json 2> gpurun_out/r3a.err
grep -h "e2e phases" gpurun_out/r3a.err
grep -h "CommitPlan\|commit() in python" gpurun_out/r3a.err | tail -24
