# streaming end-to-end loop: where the host time of the enqueue goes; wave copies merged or not
JXLB200_RUN_TIMING=1 python bench.py --no-cpu-baseline --no-also --steps 8 > gpurun_out/r2w_a.json 2> gpurun_out/r2w_a.err
grep -h "e2e phases" gpurun_out/r2w_a.err
grep -h "Run enqueue" gpurun_out/r2w_a.err | tail -12
grep -h "UploadPlan" gpurun_out/r2w_a.err | tail -6
JXLB200_NO_COALESCE=1 python bench.py --no-cpu-baseline --no-also --steps 8 > gpurun_out/r2w_b.json 2> gpurun_out/r2w_b.err
grep -h "e2e phases" gpurun_out/r2w_b.err
python - <<PY
import json
for n in "ab":
    j = json.loads(open("gpurun_out/r2w_%s.json" % n).read().strip().splitlines()[-1])
    print(n, "value %.0f e2e %.0f ms/step %.1f" % (j["value"], j["e2e"]["value"], j["ms_per_step"]))
PY
