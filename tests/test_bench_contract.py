"""bench.py's reference arm (runs without a GPU): one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-seconds", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "Mpx/s" and j["higher_is_better"] is True and j["value"] > 0
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and "sample" in j["cpu_baseline"]
    assert j["e2e"] == {"value": j["value"], "unit": "Mpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"] and j["steps"] == 1 and j["warmup"] == 0


def test_numa_binding_helper_without_sysfs():
    # (one rank: nothing is bound, the helper reports the CPUs the process may use)
    sys.path.insert(0, ROOT)
    import bench

    class _T:
        class cuda:
            @staticmethod
            def device_count():
                return 0

    info = bench.bind_rank_to_numa(_T, 0, 1)
    assert info["cpus"] == info["cpus_total"] == len(os.sched_getaffinity(0)) and info["ranks_on_node"] == 1
