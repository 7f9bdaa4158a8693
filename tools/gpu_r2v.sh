# streaming end-to-end loop: planning threads per handle
for t in 2 1 4; do
python bench.py --no-cpu-baseline --no-also --steps 8 --e2e-plan-threads $t > gpurun_out/r2v_t$t.json 2> gpurun_out/r2v_t$t.err
grep -h "e2e phases" gpurun_out/r2v_t$t.err
python - <<PY
import json
j = json.loads(open("gpurun_out/r2v_t$t.json").read().strip().splitlines()[-1])
print("threads $t: value %.0f e2e %.0f ms/step %.1f" % (j["value"], j["e2e"]["value"], j["ms_per_step"]))
PY
done
