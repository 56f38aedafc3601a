#!/usr/bin/env python
"""Generation at scale, second variant (SURVEY 8e "measure both"): one rank's share of fixed_probability(p) at n x n with
the counter-based generator, whose (row, block) units need nothing from the rest of the matrix — a rank generates exactly
its own columns, no stream is replayed and nothing is shipped.  Beside it (--stream) the same share with the reference's
sequential stream, which every rank has to replay in full (tools/c2_sharded.py is the 8-GPU run of that variant).

    python tools/bench_generator_share.py [--size 1000000] [--prob 0.1] [--world 8] [--rank 3] [--stream]

One JSON line; checks the share's structure (ascending rows inside the column range, mean degree within 5 sigma)."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import spice2_b200 as sp  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", dest="n", type=int, default=1000000)
    ap.add_argument("--prob", dest="p", type=float, default=0.1)
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--rank", type=int, default=3)
    ap.add_argument("--stream", action="store_true", help="the reference's sequential stream (bit-exact) instead")
    a = ap.parse_args()
    lo, hi = a.n * a.rank // a.world, a.n * (a.rank + 1) // a.world
    sp.generate_fixed_probability(2000, 2000, 0.1, copy=False, fast=not a.stream)  # warm-up
    t0 = time.perf_counter()
    adj = sp.Adjacency(a.n, a.n, a.p, (1337,), col_lo=lo, col_hi=hi, fast=not a.stream)
    wall = time.perf_counter() - t0
    off = adj.offsets()
    deg = np.diff(off)
    mean, sd = (hi - lo) * a.p, ((hi - lo) * a.p * (1 - a.p)) ** 0.5
    rows = adj.rows(0, 64)
    ok = bool(off[0] == 0 and off[-1] == adj.edges and abs(deg.mean() - mean) < 5 * sd / a.n ** 0.5 and rows.min() >= 0 and rows.max() < hi - lo)
    for r in range(64):
        seg = rows[off[r] - off[0]: off[r + 1] - off[0]]
        ok = ok and bool(np.all(np.diff(seg) > 0))
    print(json.dumps({"generator": "reference stream" if a.stream else "counter-based", "n": a.n, "p": a.p, "rank": a.rank, "world": a.world,
                      "columns": [lo, hi], "edges": int(adj.edges), "device_ms": adj.total_ms, "wall_ms": wall * 1e3,
                      "edges_per_s": adj.edges / (adj.total_ms * 1e-3), "write_GBps": 4.0 * adj.edges / (adj.total_ms * 1e-3) / 1e9,
                      "whole_matrix_edges_per_s_at_world_ranks": a.world * adj.edges / (adj.total_ms * 1e-3), "structure_ok": ok}))


if __name__ == "__main__":
    main()
