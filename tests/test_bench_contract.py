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


def test_round2_lines_are_verified_carry_the_cpu_arm_at_every_n_and_share_a_config():
    """Round 2: every recorded line says whether what it timed was checked against
    the oracle, N > 1 lines carry cpu_baseline too, and the reference arm's config
    equals the CUDA arm's (the driver compares them)."""
    for n, rel, ref in ((1, "profiles/r22/bench_n1.json", "profiles/r22/bench_reference_n1.json"),
                        (4, "profiles/r26/bench_n4.json", "profiles/r26/bench_reference_n4.json"),
                        (8, "profiles/r27/bench_n8.json", "profiles/r27/bench_reference_n8.json")):
        d, r = _line(rel), _line(ref)
        assert d["n_gpus"] == n == r["n_gpus"]
        assert d["config"] == r["config"], (n, d["config"], r["config"])
        assert d["cpu_baseline"] is not None and d["cpu_baseline"]["cores"] >= 1
        assert "verification" in d and d["verification"]["elements_checked"] >= 1_000_000
        assert d["e2e"]["verified"]["ok"] is True and d["e2e"]["verified"]["mismatches"] == 0
        assert d["staged_path"]["verified"]["ok"] is True
        assert d["path_resolved"] in ("fused", "nvls", "staged")
        if d["path_resolved"] == "nvls":
            assert d["unicast_ab"]["verified"]["mismatches"] == 0     # the bit-exact A/B beside it
            assert d["roofline"]["bound"] == "nvlink"
    assert _line("profiles/r27/bench_n8.json")["verified"] is True
    assert _line("profiles/r22/bench_n1.json")["verified"] is True


def test_round2_scaling_is_monotone_and_n8_beats_the_unicast_round():
    v = {n: _line(rel) for n, rel in ((1, "profiles/r22/bench_n1.json"),
                                      (4, "profiles/r26/bench_n4.json"),
                                      (8, "profiles/r27/bench_n8.json"))}
    assert v[1]["value"] < v[4]["value"] < v[8]["value"]
    assert v[8]["ms_per_step"] < v[8]["unicast_ab"]["ms_per_step"]


def test_synthetic_gradient_hash_is_the_same_in_numpy_and_torch():
    """bench.py's verification re-derives any worker's gradient from (index, seed) on
    the host; the values the GPU consumed come from the torch twin of that hash."""
    import importlib.util
    import sys

    import numpy as np
    import torch
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    sys.modules["bench_mod"] = bench
    spec.loader.exec_module(bench)
    for seed in (100, 117, 228):
        t = torch.empty(100_003, dtype=torch.float32)
        bench.synth_torch(t, seed)
        idx = np.arange(0, 100_003, 7, dtype=np.int64)
        want = bench.synth_np(idx, seed)
        assert np.array_equal(t.numpy()[idx].view(np.uint32), want.view(np.uint32))
    b = torch.empty(4096, dtype=torch.bfloat16)
    bench.synth_torch(b, 5)
    from oracle import ps_oracle as o
    want = o.f32_to_bf16(bench.synth_np(np.arange(4096, dtype=np.int64), 5))
    assert np.array_equal(b.view(torch.int16).numpy().view(np.uint16), want)
