python bench.py --no-cpu-baseline --no-also > gpurun_out/r3h.json 2> gpurun_out/r3h.err
grep -h "e2e phases" gpurun_out/r3h.err
python - <<PY
import json
j = json.loads(open("gpurun_out/r3h.json").read().strip().splitlines()[-1])
print("value %.0f e2e %.0f ms/step %.1f" % (j["value"], j["e2e"]["value"], j["ms_per_step"]), j["e2e"]["includes"][-80:])
PY
