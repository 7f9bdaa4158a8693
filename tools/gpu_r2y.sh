# streaming end-to-end loop with pinned table staging and pinned status words

JXLB200_RUN_TIMING=1 python bench.py --no-cpu-baseline --no-also --steps 8 > gpurun_out/r2y_a.json 2> gpurun_out/r2y_a.err
grep -h "e2e phases" gpurun_out/r2y_a.err
grep -h "RunToHost" gpurun_out/r2y_a.err | tail -8
grep -h "UploadPlan\|CommitPlan" gpurun_out/r2y_a.err | tail -8
python - <<PY
import json
for n in "a":
    j = json.loads(open("gpurun_out/r2x_%s.json" % n).read().strip().splitlines()[-1])
    print(n, "value %.0f e2e %.0f ms/step %.1f" % (j["value"], j["e2e"]["value"], j["ms_per_step"]))
PY
