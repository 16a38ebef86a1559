"""The JSON line `bench.py` prints is a contract with the driver: required keys of both arms.  CPU only -- the B200 arm is checked
on the committed final line of the round (profiles/), the reference arm by running it on a one-step sample."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e"}


def _last_json_line(text):
    lines = [l for l in text.splitlines() if l.startswith("{")]
    assert len(lines) == 1, "stdout must carry exactly one JSON line"
    return json.loads(lines[0])


def test_b200_arm_line_has_the_contract_keys():
    baseline = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    for name in ("r06_bench_line_b200.json", "r06_bench_line_b200_8gpu.json"):
        d = _last_json_line(open(os.path.join(ROOT, "profiles", name)).read().split("NCCL version")[-1])
        assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
        assert d["metric"] == baseline["metric"]
        assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
        assert "workload" in d["config"] and "model" not in d["config"]
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] != d["value"]
        assert d["gpu_launches"] > 0
        assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
        assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"])
        r = d["roofline"]
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] in ("hbm", "tensor")
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.0 < r["frac"] <= 1.2
        assert abs(d["value"] - d["n_gpus"] * d["config"]["batch_per_gpu"] / (1000 * d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    d1 = _last_json_line(open(os.path.join(ROOT, "profiles", "r06_bench_line_b200.json")).read())
    c = d1["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] in ("port", "reference") and c["cores"] >= 1


def test_reference_arm_prints_one_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = _last_json_line(r.stdout)
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["value"] > 0 and d["dtype"] == "f32"
