"""GPU parity: fixed_probability generation through the C ABI vs the oracle / golden vectors."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

hx = lambda v: f"{int(v):016x}"


@pytest.fixture(scope="module")
def sp():
    import spice2_b200 as sp

    assert sp.lib().spice_device_check(0) == 0, "needs a B200 (sm_100)"
    return sp


@pytest.mark.parametrize("idx", range(12))
def test_generate_matches_golden_and_oracle(sp, orc, golden, idx):
    g = golden["fixed_probability"][idx]
    r = sp.generate_fixed_probability(g["src"], g["dst"], g["p"], (1337,), g["increments"])
    assert r["edges"] == g["edges"]
    assert r["draws"] == g["edges"] + g["src"]
    assert hx(orc.fnv(r["offsets"])) == g["fnv_offsets"]
    assert hx(orc.fnv(r["neighbors"])) == g["fnv_neighbors"]
    o = orc.fixed_probability(g["src"], g["dst"], g["p"], orc.seed_seq([1337], g["increments"]))
    assert np.array_equal(r["offsets"], o["offsets"])
    assert np.array_equal(r["neighbors"], o["neighbors"])


def test_generate_1e5_golden(sp, orc, golden):
    """BASELINE config 2 at 1e5 x 1e5 (999,991,208 edges): golden hashes from the compiled reference."""
    g = golden["fixed_probability"][12]
    r = sp.generate_fixed_probability(g["src"], g["dst"], g["p"], (1337,), 0)
    assert r["edges"] == g["edges"]
    assert hx(orc.fnv(r["offsets"])) == g["fnv_offsets"]
    assert hx(orc.fnv(r["neighbors"])) == g["fnv_neighbors"]


def test_generate_degenerate(sp):
    for (a, b, p) in [(0, 10, 0.5), (10, 0, 0.5), (10, 10, 0.0)]:
        r = sp.generate_fixed_probability(a, b, p)
        assert r["edges"] == 0 and not r["offsets"].any()


def test_generate_random_shapes_vs_oracle(sp, orc):
    rng = np.random.default_rng(11)
    for _ in range(10):
        s, d = int(rng.integers(1, 3000)), int(rng.integers(1, 20000))
        p = float(rng.choice([0.002, 0.02, 0.1, 0.33, 0.75, 1.0]))
        inc = int(rng.integers(0, 5))
        r = sp.generate_fixed_probability(s, d, p, (7, 9), inc)
        o = orc.fixed_probability(s, d, p, orc.seed_seq([7, 9], inc))
        assert r["edges"] == o["edges"], (s, d, p)
        assert np.array_equal(r["offsets"], o["offsets"]) and np.array_equal(r["neighbors"], o["neighbors"]), (s, d, p)


def test_generate_column_slices_tile_the_matrix(sp, orc):
    """Multi-GPU ownership: every rank keeps a column range with local indices; the slices of all
    ranks put side by side are the reference adjacency."""
    s, d, p, world = 700, 5000, 0.1, 4
    o = orc.fixed_probability(s, d, p, orc.seed_seq([1337]))
    rows = [[] for _ in range(s)]
    for r in range(world):
        lo, hi = d * r // world, d * (r + 1) // world
        part = sp.generate_fixed_probability(s, d, p, (1337,), 0, col_lo=lo, col_hi=hi)
        for i in range(s):
            seg = part["neighbors"][part["offsets"][i]: part["offsets"][i + 1]]
            assert seg.size == 0 or (seg.min() >= 0 and seg.max() < hi - lo)
            rows[i].append(seg + lo)
    flat = np.concatenate([np.concatenate(r) for r in rows])
    assert np.array_equal(flat, o["neighbors"])


def test_generate_multi_chunk(sp, orc):
    """Force several stream chunks (rows straddling chunk borders) and compare with the oracle."""
    import ctypes as C

    L = sp.lib()
    s, d, p = 3000, 4000, 0.1
    o = orc.fixed_probability(s, d, p, orc.seed_seq([5]))
    # the public entry point picks the chunk size itself; a big enough problem crosses chunks
    r = sp.generate_fixed_probability(40000, 4000, 0.1, (5,))
    o2 = orc.fixed_probability(40000, 4000, 0.1, orc.seed_seq([5]))
    assert r["edges"] == o2["edges"] and np.array_equal(r["neighbors"], o2["neighbors"])
    r1 = sp.generate_fixed_probability(s, d, p, (5,))
    assert np.array_equal(r1["neighbors"], o["neighbors"])


def test_generate_with_every_decision_left_to_the_exact_replays(sp, orc, monkeypatch):
    """SPICE_GEN_FORCE_EXACT=1 disables the interval bounds: every row end is found by the chase's exact replay and every
    row written by fp_rows_exact — the paths that otherwise decide one row end in 1e8 and one row in 100 (and every row
    of a 1e6 x 1e6 matrix, where the bounds are too wide).  Same adjacency, bit for bit; also on column slices."""
    monkeypatch.setenv("SPICE_GEN_FORCE_EXACT", "1")
    rng = np.random.default_rng(5)
    for (s, d, p) in [(300, 4000, 0.1), (50, 30000, 0.33), (1000, 700, 0.02), (40, 3000, 1.0), (20000, 900, 0.1)]:
        inc = int(rng.integers(0, 4))
        r = sp.generate_fixed_probability(s, d, p, (3,), inc)
        o = orc.fixed_probability(s, d, p, orc.seed_seq([3], inc))
        assert r["edges"] == o["edges"], (s, d, p)
        assert np.array_equal(r["offsets"], o["offsets"]) and np.array_equal(r["neighbors"], o["neighbors"]), (s, d, p)
    s, d, p = 200, 9000, 0.1
    o = orc.fixed_probability(s, d, p, orc.seed_seq([1337]))
    part = sp.generate_fixed_probability(s, d, p, (1337,), 0, col_lo=3000, col_hi=5500)
    want = []
    for i in range(s):
        row = o["neighbors"][o["offsets"][i]: o["offsets"][i + 1]]
        want.append(row[(row >= 3000) & (row < 5500)] - 3000)
    assert np.array_equal(part["neighbors"], np.concatenate(want))


def _adj_list_reference(es, ed, src_count, col_lo, col_hi):
    """adj_list::generate (topology.cpp:63-71): sort the packed (src << 32 | dst) keys, stream them into CSR."""
    keys = np.sort((es.astype(np.int64) << 32) | ed.astype(np.int64))
    s, d = keys >> 32, keys & 0xFFFFFFFF
    keep = (d >= col_lo) & (d < col_hi)
    off = np.zeros(src_count + 1, np.int64)
    np.cumsum(np.bincount(s[keep], minlength=src_count), out=off[1:])
    return off, (d[keep] - col_lo).astype(np.int32)


def test_adj_list_generate_matches_sorted_keys(sp):
    """adj_list on the GPU (radix sort + histogram) against the reference's algorithm restated with numpy: random graphs
    with multapses, empty rows, wide target ranges (the bench's dst = i < 2^31), column slices, the empty list."""
    rng = np.random.default_rng(3)
    for (n, src, dst, lo, hi) in [(0, 5, 7, 0, 7), (1, 1, 1, 0, 1), (5000, 37, 91, 0, 91), (200000, 1000, 1 << 30, 0, 1 << 30),
                                  (100000, 5000, 3000, 700, 2100), (50000, 3, 2147483646, 0, 2147483646)]:
        es = rng.integers(0, src, n).astype(np.int32)
        ed = rng.integers(0, dst, n).astype(np.int32)
        if n > 10:
            es[:5], ed[:5] = es[5:10], ed[5:10]  # multapses
        r = sp.generate_adj_list(es, ed, src, dst, col_lo=lo, col_hi=hi)
        off, nb = _adj_list_reference(es, ed, src, lo, hi)
        assert r["edges"] == nb.size and np.array_equal(r["offsets"], off) and np.array_equal(r["neighbors"], nb), (n, src, dst)
    with pytest.raises(sp.SpiceError):
        sp.generate_adj_list(np.array([0, 9], np.int32), np.array([0, 1], np.int32), 5, 5)


# ---- the counter-based generator (spice_fixed_probability_generate_fast) ---------------------------------------------------
# Not the reference's matrix (its stream is sequential): what is pinned is (1) the algorithm as csrc/generator.cu states it
# (two definitions, by p), restated here with the oracle's seed_seq::stream / xoroshiro128+ and the host libm (whose log the
# device restates bit for bit, tests/golden/libm_pins.npz), and (2) the distribution the reference's sampler targets.
def _fast_tiles_restated(orc, seed, src, dst, p, col_lo, col_hi):
    """p >= 2^-13: blocks of B targets, unit (row, block) = engines stream((row * nblocks + block) * 32 + lane); per
    iteration a lane maps two 64-bit draws to four gaps (high word first) with the integer table T[k] = floor(2^32 q^k)."""
    import math

    lq = math.log1p(-p) if p < 1 else -math.inf
    tab = [0xFFFFFFFF]
    k = 1
    while True:
        t = int(math.floor(4294967296.0 * math.exp(k * lq)))
        tab.append(t)
        if t == 0:
            break
        k += 1
    thresholds = -np.asarray(tab[1:], np.int64)  # ascending, for searchsorted: #{k >= 1: T[k] > u}
    blog = 0
    while float(1 << blog) < 512.0 / p:
        blog += 1
    B = 1 << blog
    nblocks = (dst + B - 1) // B
    offsets, nb = [0], []
    for r in range(src):
        for b in range(col_lo // B, (col_hi + B - 1) // B):
            blk0 = b * B
            bsize = min(B, dst - blk0)
            gid = r * nblocks + b
            draws = np.stack([orc.xoroshiro(orc.L.orc_seed_stream(seed, gid * 32 + lane), 256) for lane in range(32)])  # [lane, draw]
            pos, it = -1, 0
            while pos < bsize - 1:
                x = draws[:, 2 * it: 2 * it + 2]
                u = np.stack([x[:, 0] >> np.uint64(32), x[:, 0] & np.uint64(0xFFFFFFFF), x[:, 1] >> np.uint64(32),
                              x[:, 1] & np.uint64(0xFFFFFFFF)], axis=1).astype(np.int64).reshape(-1)  # lane-major, then the four gaps
                gaps = 1 + np.searchsorted(thresholds, -u, side="left")  # #{k: T[k] > u} = #{k: -T[k] < -u}
                t = pos + np.cumsum(gaps)
                keep = t[(t < bsize) & (t + blk0 >= col_lo) & (t + blk0 < col_hi)]
                nb.extend((keep + blk0 - col_lo).tolist())
                pos = int(t[-1])
                it += 1
        offsets.append(len(nb))
    return np.asarray(offsets, np.int64), np.asarray(nb, np.int32)


def _fast_rows_restated(orc, seed, src, dst, p, col_lo, col_hi):
    """p < 2^-13: a warp per row, engines stream(row * 32 + lane), gap = 1 + floor(log(u) / log(1 - p)) in double."""
    import math

    inv = 1.0 / math.log1p(-p) if p < 1 else -0.0
    offsets, nb = [0], []
    for r in range(src):
        draws = [orc.xoroshiro(orc.L.orc_seed_stream(seed, r * 32 + lane), 64) for lane in range(32)]
        base, it = -1, 0
        while base < dst - 1:
            if it == len(draws[0]):
                draws = [orc.xoroshiro(orc.L.orc_seed_stream(seed, r * 32 + lane), 2 * it) for lane in range(32)]
            for lane in range(32):
                u = float((int(draws[lane][it]) >> 11) + 1) * 2.0 ** -53
                g = math.log(u) * inv
                base += 1 + (int(g) if g < 4.0e9 else 4000000000)
                if base < dst and col_lo <= base < col_hi:
                    nb.append(base - col_lo)
            it += 1
        offsets.append(len(nb))
    return np.asarray(offsets, np.int64), np.asarray(nb, np.int32)


def test_fast_generator_matches_its_restatement(sp, orc, monkeypatch):
    cases = [(40, 700, 0.1, 0, 700, (1337,)), (25, 30000, 0.02, 0, 30000, (7, 9)), (30, 5000, 0.75, 1000, 3200, (5,)),
             (9, 2100, 1.0, 0, 2100, (1,)), (6, 400000, 0.001, 0, 400000, (3,)), (12, 20000, 0.1, 9000, 17000, (8,)),
             (5, 300000, 0.00013, 100000, 290000, (2,)), (10, 200000, 0.006, 50000, 200000, (6,))]
    for (s, d, p, lo, hi, il) in cases:
        r = sp.generate_fixed_probability(s, d, p, il, 0, col_lo=lo, col_hi=hi, fast=True)
        off, nb = _fast_tiles_restated(orc, orc.seed_seq(list(il)), s, d, p, lo, hi)
        assert np.array_equal(r["offsets"], off), (s, d, p)
        assert np.array_equal(r["neighbors"], nb), (s, d, p)
    for (s, d, p, lo, hi, il) in [(6, 400000, 0.0001, 0, 400000, (3,)), (4, 1000000, 0.00002, 200000, 900000, (4,))]:  # the log path
        r = sp.generate_fixed_probability(s, d, p, il, 0, col_lo=lo, col_hi=hi, fast=True)
        off, nb = _fast_rows_restated(orc, orc.seed_seq(list(il)), s, d, p, lo, hi)
        assert np.array_equal(r["offsets"], off) and np.array_equal(r["neighbors"], nb), (s, d, p)
    for knob in ("SPICE_GEN_FORCE_FALLBACK", "SPICE_GEN_PIPELINED"):  # the one-buffer kernel: as the fallback, and on its own ("0")
        monkeypatch.setenv(knob, "1" if knob.endswith("FALLBACK") else "0")
        for (s, d, p, lo, hi, il) in cases[:4] + cases[-1:]:
            r = sp.generate_fixed_probability(s, d, p, il, 0, col_lo=lo, col_hi=hi, fast=True)
            off, nb = _fast_tiles_restated(orc, orc.seed_seq(list(il)), s, d, p, lo, hi)
            assert np.array_equal(r["offsets"], off) and np.array_equal(r["neighbors"], nb), (knob, s, d, p)
        monkeypatch.delenv(knob)
    monkeypatch.setenv("SPICE_GEN_FAST_LOG_PATH", "1")  # the log path at ordinary p
    for (s, d, p, lo, hi, il) in [(40, 700, 0.1, 0, 700, (1337,)), (30, 500, 0.75, 100, 320, (5,)), (9, 100, 1.0, 0, 100, (1,))]:
        r = sp.generate_fixed_probability(s, d, p, il, 0, col_lo=lo, col_hi=hi, fast=True)
        off, nb = _fast_rows_restated(orc, orc.seed_seq(list(il)), s, d, p, lo, hi)
        assert np.array_equal(r["offsets"], off) and np.array_equal(r["neighbors"], nb), (s, d, p)


def test_fast_generator_distribution_and_structure(sp):
    """Rows strictly ascending (no multapses), degrees ~ Binomial(dst, p), every pair equally likely, deterministic per
    seed; the reference's sampler has the same mean degree (topology.cpp:80-112) and clips rows at mean + 3 sigma."""
    s, d, p = 4000, 20000, 0.1
    r = sp.generate_fixed_probability(s, d, p, (1337,), fast=True)
    off, nb = r["offsets"], r["neighbors"]
    assert off[0] == 0 and off[-1] == r["edges"] == nb.size
    deg = np.diff(off)
    inner = np.ones(nb.size, bool)
    inner[off[1:-1][deg[1:] > 0]] = False  # first entry of every row but the first
    inner[0] = False
    assert (np.diff(nb.astype(np.int64))[inner[1:]] > 0).all() and nb.min() >= 0 and nb.max() < d
    mean, sd = d * p, (d * p * (1 - p)) ** 0.5
    assert abs(deg.mean() - mean) < 4 * sd / s ** 0.5
    assert 0.9 * sd < deg.std() < 1.1 * sd
    col = np.bincount(nb, minlength=d)  # in-degrees ~ Binomial(src, p): columns are not favoured by position
    assert abs(col[: d // 2].mean() - col[d // 2:].mean()) < 6 * (s * p * (1 - p)) ** 0.5 / (d / 2) ** 0.5
    assert 0.9 < col.std() / (s * p * (1 - p)) ** 0.5 < 1.1
    gaps = np.diff(nb.astype(np.int64))[inner[1:]]  # geometric: P(gap = 1) = p, mean 1/p
    assert abs((gaps == 1).mean() - p) < 0.002 and abs(gaps.mean() - 1 / p) < 0.05
    again = sp.generate_fixed_probability(s, d, p, (1337,), fast=True)
    other = sp.generate_fixed_probability(s, d, p, (1338,), fast=True)
    assert np.array_equal(again["neighbors"], nb) and not np.array_equal(other["neighbors"][:1000], nb[:1000])
    for (a, b, q) in [(0, 10, 0.5), (10, 0, 0.5), (10, 10, 0.0)]:
        z = sp.generate_fixed_probability(a, b, q, fast=True)
        assert z["edges"] == 0 and not z["offsets"].any()


def test_fast_generator_column_slices_tile_the_matrix(sp):
    s, d, p, world = 300, 9000, 0.05, 4
    full = sp.generate_fixed_probability(s, d, p, (11,), fast=True)
    rows = [[] for _ in range(s)]
    for k in range(world):
        lo, hi = d * k // world, d * (k + 1) // world
        part = sp.generate_fixed_probability(s, d, p, (11,), col_lo=lo, col_hi=hi, fast=True)
        for i in range(s):
            rows[i].append(part["neighbors"][part["offsets"][i]: part["offsets"][i + 1]] + lo)
    assert np.array_equal(np.concatenate([np.concatenate(x) for x in rows]), full["neighbors"])
