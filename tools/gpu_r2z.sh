# streaming end-to-end loop: sleeping waits; with and without the malloc thresholds, and spinning waits for comparison
run() {
  name=$1; shift
  env "$@" python bench.py --no-cpu-baseline --no-also --steps 8 $EXTRA > gpurun_out/r2z_$name.json 2> gpurun_out/r2z_$name.err
  grep -h "e2e phases" gpurun_out/r2z_$name.err
  python - <<PY
import json
j = json.loads(open("gpurun_out/r2z_$name.json").read().strip().splitlines()[-1])
print("$name: value %.0f e2e %.0f ms/step %.1f" % (j["value"], j["e2e"]["value"], j["ms_per_step"]))
PY
}
EXTRA="" run sleep_mallopt A=1
EXTRA="--no-mallopt" run sleep_plain A=1
EXTRA="" run spin_mallopt JXLB200_SPIN_WAIT=1
EXTRA="--e2e-plan-threads 4" run sleep_mallopt_t4 A=1
