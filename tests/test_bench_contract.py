"""The bench line's contract, checked on the committed line of the round's last build (profiles/bench_r02_final.json, written by
`python bench.py --steps 20 --warmup 5` on a B200): every key the driver and the judge read is there and consistent."""
import json
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_committed_bench_line_has_the_contract_keys():
    d = json.loads((ROOT / "profiles" / "bench_r02_final.json").read_text().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 20 and d["warmup"] == 5 and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert d["gpu_launches"] > 0 and d["value"] > 0
    e = d["e2e"]
    assert e["value"] > 0 and e["unit"] == d["unit"] and e["d2h_bytes_per_step"] > 0 and e["h2d_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
    assert r["traffic"] is None or r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] == 1 and c["value"] > 0 and c["unit"] == d["unit"] and c["sample"]
    k = d["clocks"]
    assert k["sm_mhz"] > 0 and k["sm_max_mhz"] >= k["sm_mhz"] and not set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["parity_check"]["matches_reference"] is True
    # the timed region: K bench steps of 150 time steps; value = events / time
    assert abs(d["ms_per_step"] * d["steps"] - d["config"]["timed_time_steps"] * d["ms_per_time_step"]) < 1e-6 * d["ms_per_step"] * d["steps"] + 1e-9
