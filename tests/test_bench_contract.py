"""The JSON lines bench.py printed on the B200 (committed under profiles/) carry every key the driver's contract
names; the metric / unit / workload are BASELINE.json's.  CPU-only: reads the committed lines."""
from __future__ import annotations

import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def test_b200_arm_line():
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    for name, n in (("r01_bench.json", 1), ("r01_bench_n2.json", 2), ("r01_bench_n8.json", 8)):
        d = load(name)
        assert d["metric"] == base["metric"] and d["unit"] == "MS/s" and d["n_gpus"] == n
        for k in ("value", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                  "config", "roofline", "e2e", "gpu_launches", "clocks"):
            assert k in d, (name, k)
        assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
        assert d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["data"] == "synthetic" and d["dtype"] == "f32"
        assert "workload" in d["config"] and "4096" in d["config"]["workload"] and "l2" in d["config"]
        r = d["roofline"]
        assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s"
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1.2
        e = d["e2e"]
        assert e["unit"] == "MS/s" and e["h2d_bytes_per_step"] == 4096 * 65472 * 2 and e["d2h_bytes_per_step"] > 0
        assert 0 < e["value"] < d["value"]                  # PCIe-bound: below the device-resident figure
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        if n == 1:
            c = d["cpu_baseline"]
            assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
            assert r["traffic"] is None or r["traffic"] > 0


def test_reference_arm_line():
    d = load("r01_bench_reference.json")
    b = load("r01_bench.json")
    assert d["impl"] == "reference" and d["metric"] == b["metric"] and d["unit"] == b["unit"]
    assert d["config"]["workload"] == b["config"]["workload"] and d["higher_is_better"] is True
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "reference"
    assert d["gpu_launches"] == 0
