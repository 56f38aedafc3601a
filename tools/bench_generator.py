#!/usr/bin/env python
"""BASELINE config 2 (bench/connectivity.cpp:24-35): fixed_probability(0.1) synapse generation at
n x n, seed {1337}, on one B200.  Prints one JSON line per size: edges, device time, edges/s and the
fraction of the write roofline (4 B per edge + 8 B per row against the measured copy bandwidth)."""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import spice2_b200 as sp  # noqa: E402

peak = 6548.5
f = ROOT / "MEASURED_PEAKS.json"
if f.exists():
    peak = float(json.loads(f.read_text())["hbm_gbs"])
fast = "--fast" in sys.argv  # the counter-based generator (not the reference's matrix)
sizes = [int(x) for x in sys.argv[1:] if x != "--fast"] or [10000, 31623, 100000]
sp.generate_fixed_probability(2000, 2000, 0.1, copy=False)  # warm-up (module load, log table upload)
for n in sizes:
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        r = sp.generate_fixed_probability(n, n, 0.1, (1337,), copy=False, fast=fast)
        wall = time.perf_counter() - t0
        if best is None or r["total_ms"] < best["total_ms"]:
            best = dict(r, wall_ms=wall * 1e3)
    gbs = (4.0 * best["edges"] + 8.0 * n) / (best["total_ms"] * 1e-3) / 1e9
    print(json.dumps({"generator": "counter-based" if fast else "reference stream", "n": n, "p": 0.1, "edges": best["edges"], "device_ms": best["total_ms"], "rows_kernel_ms": best.get("rows_ms"),
                      "wall_ms": best["wall_ms"], "edges_per_s": best["edges"] / (best["total_ms"] * 1e-3),
                      "write_GBps": gbs, "frac_of_measured_hbm": gbs / peak}))
