# record runs after the end-to-end work: all GPU tests, smoke, the default bench line
python -m pytest tests -m gpu -x -q > gpurun_out/r3c_pytest.log 2>&1; tail -3 gpurun_out/r3c_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/r3c_default.json 2> gpurun_out/r3c_default.err
grep -h "e2e phases" gpurun_out/r3c_default.err
python - <<PY
import json
j = json.loads(open("gpurun_out/r3c_default.json").read().strip().splitlines()[-1])
print("value %.0f e2e %.0f ms/step %.1f cpu %.0f" % (j["value"], j["e2e"]["value"], j["ms_per_step"], j["cpu_baseline"]["value"]))
print({k: (v.get("value"), (v.get("e2e") or {}).get("value")) for k, v in j.get("also", {}).items() if isinstance(v, dict)})
PY
