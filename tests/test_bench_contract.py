"""The JSON line bench.py printed on the B200 (committed under profiles/) carries
every key of the measurement contract; guards against the contract and the
recorded evidence drifting apart."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(rel):
    txt = open(os.path.join(ROOT, rel)).read()
    return json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])


def test_recorded_b200_line_has_the_contract_keys():
    d = _line("profiles/r20/bench.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
              "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "gpu_launches", "clocks", "roofline", "e2e", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "ps_push_pull_GBps" and d["unit"] == "GB/s" and d["n_gpus"] == 1
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor", "nvlink") and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] <= 1.05 * r["algorithmic_bytes_per_launch"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and "sample" in c
    clk = d["clocks"]
    assert not set(clk["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert clk["sm_mhz"] >= 0.9 * clk["sm_max_mhz"]


def test_recorded_reference_arm_line_has_the_contract_keys():
    d = _line("profiles/r20/bench_reference.json")
    assert d["impl"] == "reference" and d["metric"] == "ps_push_pull_GBps"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "port"


def test_recorded_scaling_lines_are_weak_scaling_of_the_same_metric():
    vals = {}
    for n, rel in ((1, "profiles/r20/bench.json"), (2, "profiles/r16/bench_n2.json"),
                   (4, "profiles/r16/bench_n4.json"), (8, "profiles/r11/bench_n8.json")):
        d = _line(rel)
        assert d["n_gpus"] == n and d["metric"] == "ps_push_pull_GBps" and d["scaling"] == "weak"
        vals[n] = d["value"]
    assert vals[1] < vals[2] < vals[4] < vals[8]
