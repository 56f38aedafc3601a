"""Pins the CPU restatement (oracle/spice_oracle.c) against golden vectors generated from the
compiled reference (tests/golden/make_golden.py) and, when oracle/_ref is present, against the
compiled reference directly.  CPU only."""
import numpy as np
import pytest

from oracle_lib import (LIF_BRUNEL, POISSON, REFBUILD, STRICT, brunel_oracle, flatten_raster, run_raster,
                        vogels_oracle)

hx = lambda v: f"{int(v):016x}"


def test_seed_seq_golden(orc, golden):
    for k, want in golden["seed_1337"].items():
        assert [hx(x) for x in orc.seed_seq([1337], int(k)).tup()] == want
    for il, want in golden["seed_il"].items():
        assert [hx(x) for x in orc.seed_seq([int(x) for x in il.split(",")]).tup()] == want


def test_xoroshiro_golden(orc, golden):
    assert [hx(x) for x in orc.xoroshiro(orc.seed_seq([1337]), 8)] == golden["xoroshiro_1337_first8"]
    assert hx(orc.fnv(orc.xoroshiro(orc.seed_seq([1337], 3), 100000))) == golden["xoroshiro_1337_inc3_fnv_1e5"]


def test_kahan_dt_golden(orc, golden):
    g = golden["kahan_dt_1e-4"]
    dt = orc.kahan_dt(np.float32(1e-4), 30000)
    assert [hx(x) for x in dt[:8].view(np.uint32)] == g["first8"]
    assert hx(dt[299:300].view(np.uint32)[0]) == g["i299"]
    assert hx(orc.fnv(dt)) == g["fnv_30000"]
    assert len(np.unique(dt[:300])) == g["distinct_300"]


@pytest.mark.parametrize("idx", range(12))
def test_fixed_probability_golden(orc, golden, idx):
    g = golden["fixed_probability"][idx]
    r = orc.fixed_probability(g["src"], g["dst"], g["p"], orc.seed_seq([1337], g["increments"]))
    assert r["capacity"] == g["capacity"]
    assert r["edges"] == g["edges"]
    deg = np.diff(r["offsets"])
    assert (int(deg.min()), int(deg.max())) == (g["deg_min"], g["deg_max"])
    assert [int(x) for x in r["neighbors"][:8]] == g["row0"]
    assert hx(orc.fnv(r["offsets"])) == g["fnv_offsets"]
    assert hx(orc.fnv(r["neighbors"])) == g["fnv_neighbors"]
    # structural properties (topology.cpp:99-107): rows strictly ascending, in range, capped
    assert r["draws"] == r["edges"] + g["src"]
    assert deg.max() <= orc.max_degree(g["dst"], g["p"])
    if r["edges"]:
        nb, off = r["neighbors"], r["offsets"]
        inc = np.diff(nb) > 0
        row_start = np.zeros(len(nb), bool)
        row_start[off[:-1][deg > 0]] = True
        assert np.all(inc | row_start[1:])
        assert nb.min() >= 0 and nb.max() < g["dst"]


@pytest.mark.slow
def test_fixed_probability_1e5_golden(orc, golden):
    g = golden["fixed_probability"][12]
    assert (g["src"], g["dst"]) == (100000, 100000)
    r = orc.fixed_probability(g["src"], g["dst"], g["p"], orc.seed_seq([1337]), want_neighbors=False)
    assert r["edges"] == g["edges"] == 999991208
    assert hx(orc.fnv(r["offsets"])) == g["fnv_offsets"]


def test_fixed_probability_degenerate(orc):
    s = orc.seed_seq([1337])
    for (a, b, p) in [(0, 10, 0.5), (10, 0, 0.5), (10, 10, 0.0)]:
        assert orc.fixed_probability(a, b, p, s)["edges"] == 0


def _check_sample(orc, g, net, pops, state_pops):
    rows, cnt = run_raster(net, pops, g["steps"])
    assert [int(x) for x in cnt.sum(0)] == g["totals"]
    assert hx(orc.fnv(cnt)) == g["fnv_counts"]
    assert hx(orc.fnv(flatten_raster(rows))) == g["fnv_ids"]
    assert hx(orc.fnv(net.neurons(state_pops[0]))) == g["fnv_state_E"]
    assert hx(orc.fnv(net.neurons(state_pops[1]))) == g["fnv_state_I"]


def test_brunel_300_golden(orc, golden):
    net, pops = brunel_oracle(orc)
    _check_sample(orc, golden["samples"]["brunel_300_strict"], net, pops, (1, 2))
    # the raster of the reference's own build is identical over the sample's 300 steps
    for k in ("totals", "fnv_counts", "fnv_ids"):
        assert golden["samples"]["brunel_300_fast"][k] == golden["samples"]["brunel_300_strict"][k]


def test_brunel_3000_refbuild_golden(orc, golden):
    net, pops = brunel_oracle(orc, flavour=REFBUILD)
    _check_sample(orc, golden["samples"]["brunel_3000_fast"], net, pops, (1, 2))


def test_brunel_3000_strict_golden(orc, golden):
    net, pops = brunel_oracle(orc, flavour=STRICT)
    _check_sample(orc, golden["samples"]["brunel_3000_strict"], net, pops, (1, 2))


def test_brunel_small_golden(orc, golden):
    net, pops = brunel_oracle(orc, N=3010, p=0.07, w_exc=np.float32(2.0 / 300), w_inh=np.float32(-10.0 / 300),
                              delay=7e-4, seed=(42,))
    _check_sample(orc, golden["samples"]["brunel_small_strict"], net, pops, (1, 2))


def test_vogels_1500_golden(orc, golden):
    net, pops = vogels_oracle(orc)
    _check_sample(orc, golden["samples"]["vogels_1500_strict"], net, pops, (0, 1))


@pytest.mark.slow
def test_brunel_plus_300_golden(orc, golden):
    net, pops = brunel_oracle(orc, plastic=True)
    _check_sample(orc, golden["samples"]["brunel_plus_300_strict"], net, pops, (1, 2))


def test_spikes_age_precondition(orc):
    net, (P, E, I) = brunel_oracle(orc, N=200)
    with pytest.raises(ValueError):
        net.spikes(P, 0)  # no step run yet (neuron_population.h:148)
    net.step()
    net.spikes(P, 0)
    with pytest.raises(ValueError):
        net.spikes(P, 1)


def test_connect_delay_precondition(orc):
    net = orc.net(np.float32(1e-4), np.float32(15e-4))
    a = net.add_population(POISSON, 10)
    b = net.add_population(LIF_BRUNEL, 10)
    with pytest.raises(ValueError):
        net.connect(a, b, 0.1, np.float32(16e-4), 0, 0.0)  # snn.h:36-38
    with pytest.raises(ValueError):
        net.connect(a, b, 0.1, np.float32(0.0), 0, 0.0)  # snn.h:35


def test_sharded_oracle_equals_unsharded(orc):
    """Target-partitioned execution with a per-step spike exchange reproduces the
    single-instance run bit for bit."""
    kw = dict(N=1500, p=0.1, w_exc=np.float32(2.0 / 300), w_inh=np.float32(-10.0 / 300))
    full, pops = brunel_oracle(orc, **kw)
    world = 3
    shards = [brunel_oracle(orc, rank=r, world=world, **kw)[0] for r in range(world)]
    for step in range(120):
        full.step()
        for s in shards:
            s.step_update()
        for p in pops:
            ids = np.concatenate([s.local_spikes(p) for s in shards])
            assert np.array_equal(ids, full.spikes(p, 0))
            for s in shards:
                s.step_set_spikes(p, ids)
        for s in shards:
            s.step_deliver()
    for p in pops[1:]:
        assert np.array_equal(np.concatenate([s.neurons(p) for s in shards]), full.neurons(p))
    assert sum(s.events() for s in shards) == full.events()


# ---- directly against the compiled reference, when it is here --------------------------------
def test_oracle_vs_reference_fixed_probability(orc, ref_fast):
    rng = np.random.default_rng(7)
    for _ in range(12):
        s, d = int(rng.integers(1, 400)), int(rng.integers(1, 3000))
        p = float(rng.choice([0.003, 0.02, 0.1, 0.33, 0.75, 1.0]))
        inc = int(rng.integers(0, 4))
        a = orc.fixed_probability(s, d, p, orc.seed_seq([99, 5], inc))
        b = ref_fast.fixed_probability(s, d, p, [99, 5], inc)
        assert a["edges"] == b["edges"] and a["capacity"] == b["capacity"]
        assert np.array_equal(a["offsets"], b["offsets"]) and np.array_equal(a["neighbors"], b["neighbors"])


def test_oracle_vs_reference_brunel(orc, ref_strict, ref_fast):
    kw = dict(N=2500, p=0.1, w_exc=np.float32(2.0 / 500), w_inh=np.float32(-10.0 / 500), delay=11e-4)
    for shim, flavour in ((ref_strict, STRICT), (ref_fast, REFBUILD)):
        net, pops = brunel_oracle(orc, seed=(5,), flavour=flavour, **kw)
        rows, cnt = run_raster(net, pops, 500)
        r = shim.brunel(seed=5, steps=500, **kw)
        assert np.array_equal(cnt, r["counts"]) and np.array_equal(flatten_raster(rows), r["ids"])
        assert np.array_equal(net.neurons(1), r["state_E"]) and np.array_equal(net.neurons(2), r["state_I"])


def test_oracle_vs_reference_brunel_plus(orc, ref_strict):
    kw = dict(N=2500, p=0.1, w_exc=np.float32(2.0 / 500), w_inh=np.float32(-10.0 / 500))
    net, pops = brunel_oracle(orc, seed=(5,), plastic=True, **kw)
    rows, cnt = run_raster(net, pops, 200)
    r = ref_strict.brunel(seed=5, steps=200, plastic=True, **kw)
    assert np.array_equal(cnt, r["counts"]) and np.array_equal(flatten_raster(rows), r["ids"])
    assert np.array_equal(net.neurons(1), r["state_E"])


def test_oracle_vs_reference_vogels_dense(orc, ref_strict):
    """The dense Vogels network of tests/test_gpu_sim.py::test_delivery_whole_units_and_single_rounds (p = 0.5:
    whole-population volleys, rows of 1600 + 400 entries): the restatement that checks the GPU there agrees with
    the compiled reference here."""
    net, pops = vogels_oracle(orc, p=0.5)
    rows, cnt = run_raster(net, pops, 200)
    r = ref_strict.vogels(p=0.5, steps=200)
    assert cnt.max() == 3200
    assert np.array_equal(cnt, r["counts"]) and np.array_equal(flatten_raster(rows), r["ids"])
    assert np.array_equal(net.neurons(0), r["state_E"]) and np.array_equal(net.neurons(1), r["state_I"])
