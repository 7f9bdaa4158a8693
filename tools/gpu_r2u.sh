# streaming end-to-end loop (plan ahead + per-wave read-back) against the serial one, and the new GPU tests
python -m pytest tests/test_gpu_decode.py tests/test_gpu_vardct.py -m gpu -x -q > gpurun_out/r2u_pytest.log 2>&1
tail -3 gpurun_out/r2u_pytest.log
python bench.py --no-cpu-baseline --no-also > gpurun_out/r2u_stream.json 2> gpurun_out/r2u_stream.err
python bench.py --no-cpu-baseline --no-also --e2e-serial > gpurun_out/r2u_serial.json 2> gpurun_out/r2u_serial.err
grep -h "e2e" gpurun_out/r2u_stream.err gpurun_out/r2u_serial.err
python - <<'PY'
import json
for n in ("stream", "serial"):
    try:
        j = json.loads(open("gpurun_out/r2u_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "value %.0f e2e %.0f ms/step %.1f ok %s" % (j["value"], j["e2e"]["value"], j["ms_per_step"], j["config"]["golden_checksum_ok"]))
    except Exception as e:
        print(n, "failed", e)
PY
