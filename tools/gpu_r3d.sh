# warp-per-section rANS emission: encoder parity tests, encode benches
python -m pytest tests/test_encoder.py -m gpu -x -q 2>&1 | tail -3
python bench.py --workload encode4k --batch 16 --steps 3 --warmup 1 --no-cpu-baseline --no-also > gpurun_out/r3d_enc.json 2> gpurun_out/r3d_enc.err
python - <<PY
import json
j = json.loads(open("gpurun_out/r3d_enc.json").read().strip().splitlines()[-1])
print("encode: value %.0f e2e %.0f ms/step %.1f" % (j["value"], j["e2e"]["value"], j["ms_per_step"]), j["roofline"].get("kernel"), j["roofline"].get("kernel_ms"), j["config"].get("golden_checksum_ok"))
print(j.get("phases") or j["roofline"])
PY
python tools/bench_lossless_enc.py 16 3 | tail -1
