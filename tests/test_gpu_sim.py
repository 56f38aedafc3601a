"""GPU parity: the per-timestep loop through the C ABI vs the oracle and the golden vectors."""
import numpy as np
import pytest

from oracle_lib import brunel_oracle, flatten_raster, run_raster, vogels_oracle

pytestmark = pytest.mark.gpu

hx = lambda v: f"{int(v):016x}"


@pytest.fixture(scope="module")
def sp():
    import spice2_b200 as sp

    assert sp.lib().spice_device_check(0) == 0, "needs a B200 (sm_100)"
    return sp


def gpu_raster(net, steps, batch):
    """Run `steps` steps in calls of `batch` and return (counts[steps,npop], flat ids)."""
    net.raster_enable(True)
    done = 0
    while done < steps:
        n = min(batch, steps - done)
        net.step(n)
        done += n
    return net.raster_read()


def test_brunel_300_golden(sp, orc, golden):
    from spice2_b200.samples import brunel

    g = golden["samples"]["brunel_300_strict"]
    net, (P, E, I) = brunel()
    counts, ids = gpu_raster(net, 300, 300)
    assert [int(x) for x in counts.sum(0)] == g["totals"]
    assert hx(orc.fnv(counts)) == g["fnv_counts"]
    assert hx(orc.fnv(ids)) == g["fnv_ids"]
    assert hx(orc.fnv(E.get_neurons())) == g["fnv_state_E"]
    assert hx(orc.fnv(I.get_neurons())) == g["fnv_state_I"]
    # the raster of the reference's own (-ffast-math) build is the same over the sample's 300 steps
    assert golden["samples"]["brunel_300_fast"]["fnv_ids"] == g["fnv_ids"]


def test_brunel_adjacency_matches_oracle(sp, orc):
    from spice2_b200.samples import brunel

    net, pops = brunel(N=3000)
    onet, _ = brunel_oracle(orc, N=3000)
    sizes = [1500, 1500, 1200, 1200, 300, 300]
    for ci in range(6):
        off, nb = net.connection_csr(ci)
        ooff, onb = onet.connection_csr(ci, sizes[ci])
        assert np.array_equal(off, ooff) and np.array_equal(nb, onb)


@pytest.mark.parametrize("batch", [1, 7, 15, 64])
def test_brunel_small_step_by_step(sp, orc, golden, batch):
    """Any partition of the steps into launch windows gives the same result; spikes(age) and
    get_neurons() agree with the oracle at every readout."""
    from spice2_b200.samples import brunel

    kw = dict(N=3010, p=0.07, w_exc=np.float32(2.0 / 300), w_inh=np.float32(-10.0 / 300), delay=7e-4, seed=(42,))
    net, pops = brunel(**kw)
    onet, opops = brunel_oracle(orc, **kw)
    steps, done = 120, 0
    while done < steps:
        n = min(batch, steps - done)
        net.step(n)
        for _ in range(n):
            onet.step()
        done += n
        for p, op in zip(pops, opops):
            for age in range(min(n, 3)):
                assert np.array_equal(p.spikes(age), onet.spikes(op, age)), (done, age)
        assert np.array_equal(pops[1].get_neurons(), onet.neurons(1))
        assert np.array_equal(pops[2].get_neurons(), onet.neurons(2))


def test_brunel_small_golden(sp, orc, golden):
    from spice2_b200.samples import brunel

    g = golden["samples"]["brunel_small_strict"]
    net, (P, E, I) = brunel(N=3010, p=0.07, w_exc=np.float32(2.0 / 300), w_inh=np.float32(-10.0 / 300), delay=7e-4, seed=(42,))
    counts, ids = gpu_raster(net, 400, 100)
    assert [int(x) for x in counts.sum(0)] == g["totals"]
    assert hx(orc.fnv(ids)) == g["fnv_ids"]
    assert hx(orc.fnv(E.get_neurons())) == g["fnv_state_E"]


def test_brunel_3000_strict_golden(sp, orc, golden):
    from spice2_b200.samples import brunel

    g = golden["samples"]["brunel_3000_strict"]
    net, (P, E, I) = brunel()
    counts, ids = gpu_raster(net, 3000, 1000)
    assert [int(x) for x in counts.sum(0)] == g["totals"]
    assert hx(orc.fnv(ids)) == g["fnv_ids"]
    assert hx(orc.fnv(E.get_neurons())) == g["fnv_state_E"]
    assert hx(orc.fnv(I.get_neurons())) == g["fnv_state_I"]


def test_vogels_1500_golden(sp, orc, golden):
    from spice2_b200.samples import vogels

    g = golden["samples"]["vogels_1500_strict"]
    net, (E, I) = vogels()
    counts, ids = gpu_raster(net, 1500, 500)
    assert [int(x) for x in counts.sum(0)] == g["totals"]
    assert hx(orc.fnv(counts)) == g["fnv_counts"]
    assert hx(orc.fnv(ids)) == g["fnv_ids"]
    assert hx(orc.fnv(E.get_neurons())) == g["fnv_state_E"]
    assert hx(orc.fnv(I.get_neurons())) == g["fnv_state_I"]


def test_synaptic_event_count(sp, orc):
    from spice2_b200.samples import brunel

    net, pops = brunel(N=3000)
    onet, opops = brunel_oracle(orc, N=3000)
    net.step(200)
    # the backend counts a spike's events when the window that emitted it is delivered; the
    # reference delivers them delay-1 steps later, so its tally catches up after 14 more steps
    for _ in range(200 + 14):
        onet.step()
    assert net.stats()["synaptic_events"] == onet.events()


def test_seed_bookkeeping(sp, golden):
    """snn.h:21-56: a stateful population consumes one seed++, a stateless one none, every connection
    one; a step consumes one more (snn.cpp:12).  Brunel: 2 + 6 increments before the first step."""
    from spice2_b200.samples import brunel

    net, _ = brunel(N=3000)
    assert [hx(x) for x in net.seed()] == golden["seed_1337"]["8"]
    net.step(15)
    net.sync()
    assert list(net.seed()) == list(sp.seed_seq([1337], 8 + 15))


def test_preconditions(sp):
    net = sp.snn(1e-4, 15e-4)
    a = net.add_population("brunel.poisson", 10)
    b = net.add_population("brunel.lif", 10)
    with pytest.raises(sp.SpiceError, match="Assertion failed"):
        net.connect("brunel.fixed_weight", a, b, sp.fixed_probability(0.1), 16e-4, weight=0.1)  # snn.h:36-38
    with pytest.raises(sp.SpiceError, match="Assertion failed"):
        net.connect("brunel.fixed_weight", a, b, sp.fixed_probability(0.1), 0.0, weight=0.1)  # snn.h:35
    with pytest.raises(sp.SpiceError, match="Assertion failed"):
        net.connect("brunel.fixed_weight", a, b, sp.fixed_probability(1.5), 1e-4, weight=0.1)  # topology.cpp:73
    with pytest.raises(sp.SpiceError):
        a.spikes(0)  # no step run yet (neuron_population.h:148)
    with pytest.raises(sp.SpiceError):
        a.get_neurons()  # stateless population (neuron_population.h:143)


def test_two_ranks_one_device(sp, orc):
    """Target-partitioned execution with the peer-store spike exchange: two rank contexts in one
    process (same GPU) reproduce the single-context run bit for bit."""
    from spice2_b200.samples import brunel

    kw = dict(N=3000, p=0.1, w_exc=np.float32(2.0 / 300), w_inh=np.float32(-10.0 / 300))
    onet, opops = brunel_oracle(orc, **kw)
    world = 2
    nets = [brunel(rank=r, world=world, **kw) for r in range(world)]
    for net, _ in nets:
        net.finalize()
    handles = [net.peer_handle() for net, _ in nets]
    for net, _ in nets:
        net.set_peers(handles)
    for chunk in range(8):
        for net, _ in nets:
            net.step(15)
        for _ in range(15):
            onet.step()
        for net, pops in nets:
            for p, op in zip(pops, opops):
                assert np.array_equal(p.spikes(0), onet.spikes(op, 0))
    for pi in (1, 2):
        got = np.concatenate([pops[pi].get_neurons() for _, pops in nets])
        assert np.array_equal(got, onet.neurons(pi))
    for _ in range(14):
        onet.step()
    assert sum(net.stats()["synaptic_events"] for net, _ in nets) == onet.events()


@pytest.mark.parametrize("shape", ["30/70", "0/25/75"])
def test_ranks_with_uneven_target_ranges(sp, orc, shape):
    """spice_set_next_partition: rank 0 owns 30 % of every population and rank 1 the rest, or three ranks of which the first
    owns nothing (ranges an in-degree balance can produce): still the reference's raster and state, bit for bit."""
    from spice2_b200.samples import brunel

    kw = dict(N=3000, p=0.1, w_exc=np.float32(2.0 / 300), w_inh=np.float32(-10.0 / 300))
    onet, opops = brunel_oracle(orc, **kw)
    world, part = (2, lambda n: [0, n * 3 // 10, n]) if shape == "30/70" else (3, lambda n: [0, 0, n // 4, n])
    nets = [brunel(rank=r, world=world, partition=part, **kw) for r in range(world)]
    assert [pops[1].range() for _, pops in nets] == ([(0, 360), (360, 1200)] if world == 2 else [(0, 0), (0, 300), (300, 1200)])
    for net, _ in nets:
        net.finalize()
    handles = [net.peer_handle() for net, _ in nets]
    for net, _ in nets:
        net.set_peers(handles)
    for chunk in range(8):
        for net, _ in nets:
            net.step(15)
        for _ in range(15):
            onet.step()
        for net, pops in nets:
            for p, op in zip(pops, opops):
                assert np.array_equal(p.spikes(0), onet.spikes(op, 0))
    for pi in (1, 2):
        got = np.concatenate([pops[pi].get_neurons() for _, pops in nets])
        assert np.array_equal(got, onet.neurons(pi))
    for _ in range(14):
        onet.step()
    assert sum(net.stats()["synaptic_events"] for net, _ in nets) == onet.events()
    with pytest.raises(sp.SpiceError, match="Assertion failed"):
        brunel(rank=0, world=2, partition=lambda n: [0, n + 1, n], **kw)


def test_in_degree_balanced_ranges_on_a_skewed_adj_list(sp):
    """Static synapse-count load balancing: an adj_list network whose in-degrees fall off like 1 / (1 + i / 40);
    spice_balance_ranges cuts the targets where the prefix sum of the in-degrees crosses r / world of the total, the two
    ranks then hold the same number of synapses within 5 % (equal widths: 4 : 1) and reproduce the one-rank run bit for bit."""
    rng = np.random.default_rng(17)
    n_src, n_dst = 1500, 2000
    deg = np.maximum((600.0 / (1.0 + np.arange(n_dst) / 40.0)).astype(np.int64), 1)
    dst = np.repeat(np.arange(n_dst, dtype=np.int32), deg)
    src = np.concatenate([rng.choice(n_src, d, replace=False) for d in deg]).astype(np.int32)
    rec_src = rng.integers(0, n_dst, 40000).astype(np.int32)
    rec_dst = np.minimum((rng.random(40000) ** 3 * n_dst).astype(np.int32), n_dst - 1)
    in_deg = np.bincount(dst, minlength=n_dst) + np.bincount(rec_dst, minlength=n_dst)
    bounds = sp.balance_ranges(in_deg, 2)
    cut = int(bounds[1])
    assert bounds[0] == 0 and bounds[2] == n_dst and 0 < cut < n_dst // 2
    share = in_deg[:cut].sum() / in_deg.sum()
    assert abs(share - 0.5) < 0.05 and in_deg[: n_dst // 2].sum() / in_deg.sum() > 0.75

    def build(rank, world):
        net = sp.snn(1e-4, 5e-4, (3,), rank=rank, world=world)
        P = net.add_population("brunel.poisson", n_src)
        E = net.add_population("brunel.lif", n_dst, bounds=None if world == 1 else bounds)
        net.connect("brunel.fixed_weight", P, E, sp.adj_list(src, dst), 5e-4, weight=np.float32(0.004))
        net.connect("brunel.fixed_weight", E, E, sp.adj_list(rec_src, rec_dst), 3e-4, weight=np.float32(-0.002))
        return net, (P, E)

    one, one_pops = build(0, 1)
    nets = [build(r, 2) for r in range(2)]
    for net, _ in nets:
        net.finalize()
    handles = [net.peer_handle() for net, _ in nets]
    for net, _ in nets:
        net.set_peers(handles)
    edges = [sum(net.connection_edges(c) for c in range(2)) for net, _ in nets]
    assert sum(edges) == len(src) + len(rec_src) and abs(edges[0] / sum(edges) - 0.5) < 0.05
    total = 0
    for chunk in range(20):
        one.step(9)
        for net, _ in nets:
            net.step(9)
        for net, pops in nets:
            for p, q in zip(pops, one_pops):
                assert np.array_equal(p.spikes(0), q.spikes(0))
        total += len(one_pops[1].spikes(0))
    assert total > 0
    got = np.concatenate([pops[1].get_neurons() for _, pops in nets])
    assert np.array_equal(got, one_pops[1].get_neurons())


# ---- plastic synapses (samples/brunel+.cpp): lazy event-driven STDP --------------------------------
# Parity contract: the plastic path calls libm (expf, pow).  expf is restated
# bit-exactly, pow(base, n) is evaluated correctly rounded, which glibc's pow is except in rare
# near-halfway cases, so the comparison below is exact in practice; the stated tolerance is what
# the test falls back to if a last-bit libm difference ever shows.
PLASTIC_V_TOL = 1e-5   # volts, on membrane potentials of ~1e-2
PLASTIC_W_TOL = 1e-9   # on weights of ~1e-4


def _close_or_equal(got, want, fields, tol):
    """Bit-exact, or (reported with a warning, never silently) within the stated tolerance."""
    if np.array_equal(got, want):
        return True
    ok = all(np.allclose(got[f], want[f], rtol=0, atol=tol) for f in fields)
    if ok:
        import warnings

        worst = max(float(np.max(np.abs(got[f].astype(np.float64) - want[f].astype(np.float64)))) for f in fields)
        warnings.warn(f"plastic path: not bit-exact, within tolerance {tol} (max abs difference {worst:.3e} over {fields})")
    return ok


def test_brunel_plus_step_by_step(sp, orc):
    from spice2_b200.samples import brunel

    kw = dict(N=3000, p=0.1, w_exc=np.float32(2.0 / 300), w_inh=np.float32(-10.0 / 300), seed=(5,), plastic=True)
    net, pops = brunel(**kw)
    onet, opops = brunel_oracle(orc, **kw)
    for step in range(200):
        net.step()
        onet.step()
        for p, op in zip(pops, opops):
            assert np.array_equal(p.spikes(0), onet.spikes(op, 0)), step
        if step % 20 == 19 or step in (63, 64, 65, 127, 128, 129):
            assert _close_or_equal(pops[1].get_neurons(), onet.neurons(1), ("V",), PLASTIC_V_TOL), step
            assert _close_or_equal(pops[2].get_neurons(), onet.neurons(2), ("V",), PLASTIC_V_TOL), step
    got, want = net.connection_synapses(2), onet.connection_synapses(2)
    assert _close_or_equal(got, want, ("W", "Zpre", "Zpost"), PLASTIC_W_TOL)
    for _ in range(14):
        onet.step()


def test_two_ranks_one_device_plastic(sp, orc):
    """brunel+ target-partitioned over two rank contexts (window = 1 step, every connection's delay ==
    max_delay, so the stateful delivery of step t reads the ring slot a peer one step ahead would
    overwrite without the slack window in the ring, runtime.cu finalize): spikes, neuron state and
    synapse state of the single-process oracle, including the 64-step flush (steps 64, 128)."""
    from spice2_b200.samples import brunel

    kw = dict(N=3000, p=0.1, w_exc=np.float32(2.0 / 300), w_inh=np.float32(-10.0 / 300), seed=(5,), plastic=True)
    onet, opops = brunel_oracle(orc, **kw)
    world = 2
    nets = [brunel(rank=r, world=world, **kw) for r in range(world)]
    for net, _ in nets:
        net.finalize()
    handles = [net.peer_handle() for net, _ in nets]
    for net, _ in nets:
        net.set_peers(handles)
    for chunk in range(10):
        for net, _ in nets:
            net.step(15)
        for _ in range(15):
            onet.step()
        for net, pops in nets:
            for age in (0, 7, 14):
                for p, op in zip(pops, opops):
                    assert np.array_equal(p.spikes(age), onet.spikes(op, age)), (chunk, age)
        for pi in (1, 2):
            got = np.concatenate([pops[pi].get_neurons() for _, pops in nets])
            assert _close_or_equal(got, onet.neurons(pi), ("V",), PLASTIC_V_TOL), chunk
    # the E->E synapses: rank r holds the columns of its targets; compare row by row with the oracle's CSR
    ooff, onb = onet.connection_csr(2, 1200)
    osyn = onet.connection_synapses(2)
    lo_hi = [pops[1].range() for _, pops in nets]
    for (net, _), (lo, hi) in zip(nets, lo_hi):
        off, nb = net.connection_csr(2)
        syn = net.connection_synapses(2)
        keep = (onb >= lo) & (onb < hi)
        assert np.array_equal(nb + lo, onb[keep])
        assert _close_or_equal(syn, osyn[keep], ("W", "Zpre", "Zpost"), PLASTIC_W_TOL)
        assert off[-1] == keep.sum()


# north star: "the fast atomic mode must match membrane potentials within a stated float tolerance and per-population
# firing rates within a stated tolerance".  SPICE_MODE_FAST applies a target's plastic events in arrival order (the
# event lists are filled with atomics) instead of the reference's (source, row) order: float sums in another order.
FAST_V_TOL = 1e-4        # volts, per neuron, on membrane potentials of ~1e-2 (threshold 0.02), for >= 99.5 % of the neurons
FAST_RATE_TOL = 0.02     # relative, per population, over the run


def test_fast_mode_tolerances(sp, orc):
    from spice2_b200.samples import brunel

    kw = dict(N=3000, p=0.1, w_exc=np.float32(2.0 / 300), w_inh=np.float32(-10.0 / 300), seed=(5,), plastic=True)
    net, pops = brunel(mode=sp.MODE_FAST, **kw)
    onet, opops = brunel_oracle(orc, **kw)
    steps = 300
    net.raster_enable(True)
    net.step(steps)
    counts, _ids = net.raster_read(steps)
    ocounts = np.zeros_like(counts)
    for s in range(steps):
        onet.step()
        ocounts[s] = [len(onet.spikes(op, 0)) for op in opops]
    got, want = counts.sum(0).astype(float), ocounts.sum(0).astype(float)
    assert np.all(np.abs(got - want) <= FAST_RATE_TOL * want + 1), (got, want)
    for pi in (1, 2):
        g, w = pops[pi].get_neurons(), onet.neurons(pi)
        close = np.abs(g["V"].astype(np.float64) - w["V"].astype(np.float64)) <= FAST_V_TOL
        assert close.mean() >= 0.995, (pi, close.mean())
    gs, ws = net.connection_synapses(2), onet.connection_synapses(2)
    assert np.mean(np.abs(gs["W"].astype(np.float64) - ws["W"].astype(np.float64)) <= 1e-7) >= 0.995


FAST_TOPOLOGY_BAND = 0.10  # the rates must lie within the spread of the reference's own seeds, widened by this much


def test_fast_topology_keeps_the_firing_rates(sp, orc):
    """fixed_probability(p, fast=True): the counter-based generator draws another matrix from the distribution the
    reference's sampler targets, so the raster differs.  At this size the rates of the E and I populations depend on the
    drawn graph by +-16 % (six reference seeds, 6000 steps), so the stated tolerance is that spread: the rates with the
    counter-based graph must lie inside the band the reference's own graphs span.  The Poisson input is the same stream
    (the connection consumes the same seed increment), spike for spike."""
    from spice2_b200.samples import brunel

    kw = dict(N=8000, p=0.1, w_exc=np.float32(2.0 / 800), w_inh=np.float32(-10.0 / 800))
    steps, skip = 1500, 300
    net, pops = brunel(fast_topology=True, seed=(1337,), **kw)
    net.raster_enable(True)
    net.step(steps)
    counts, _ids = net.raster_read(steps)
    band = []
    for seed in ((1337,), (1,), (2,), (3,), (4,), (5,)):
        onet, opops = brunel_oracle(orc, seed=seed, **kw)
        ocounts = np.zeros_like(counts)
        for s in range(steps):
            onet.step()
            ocounts[s] = [len(onet.spikes(op, 0)) for op in opops]
        if seed == (1337,):
            assert np.array_equal(counts[:, 0], ocounts[:, 0])  # the Poisson population does not depend on the graph
        band.append(ocounts[skip:].sum(0).astype(float))
    band = np.array(band)
    got = counts[skip:].sum(0).astype(float)
    assert np.all(got >= band.min(0) * (1 - FAST_TOPOLOGY_BAND)) and np.all(got <= band.max(0) * (1 + FAST_TOPOLOGY_BAND)), (got, band)
    off, nb = net.connection_csr(2)
    deg = np.diff(off)
    assert abs(deg.mean() - 320.0) < 4 * (3200 * 0.1 * 0.9) ** 0.5 / 3200 ** 0.5


def test_brunel_plus_300_golden(sp, orc, golden):
    """samples/brunel+ (N = 20000, 300 steps): raster of the compiled reference (strict flavour)."""
    from spice2_b200.samples import brunel

    g = golden["samples"]["brunel_plus_300_strict"]
    net, (P, E, I) = brunel(plastic=True)
    counts, ids = gpu_raster(net, 300, 50)
    assert [int(x) for x in counts.sum(0)] == g["totals"]
    assert hx(orc.fnv(ids)) == g["fnv_ids"]


def test_device_libm_matches_host_pins(sp):
    """exp(float) and pow(double, n) as the plastic model evaluates them on the device vs the host
    glibc values the reference would compute (tests/golden/libm_pins.npz, from the compiled reference's
    process).  expf is a bit-exact restatement; pow must be exact on the 65 values skip() can ask for."""
    from pathlib import Path

    z = np.load(Path(__file__).resolve().parent / "golden" / "libm_pins.npz")
    x = np.ascontiguousarray(z["expf_x"], np.float32)
    y = np.zeros_like(x)
    assert sp.lib().spice_selftest_libm(0, 0, x.ctypes.data, None, y.ctypes.data, len(x)) == 0
    assert np.array_equal(y.view(np.uint32), z["expf_y"].astype(np.float32).view(np.uint32))
    px = np.ascontiguousarray(z["pow_x"], np.float64)
    pn = np.ascontiguousarray(z["pow_n"], np.int64)
    py = np.zeros_like(px)
    assert sp.lib().spice_selftest_libm(0, 1, px.ctypes.data, pn.ctypes.data, py.ctypes.data, len(px)) == 0
    assert np.array_equal(py.view(np.uint64), z["pow_y"].astype(np.float64).view(np.uint64))


def test_sink_partial_reads_and_wraparound(sp, orc, monkeypatch):
    """The spike sink (spice_raster_*): readouts of any length, issued while later batches are still
    queued, return the same lists as one big readout; the host ring wraps around."""
    from spice2_b200.samples import brunel

    kw = dict(N=4000, p=0.1, w_exc=np.float32(2.0 / 400), w_inh=np.float32(-10.0 / 400))
    net, _ = brunel(**kw)
    counts0, ids0 = gpu_raster(net, 600, 600)
    monkeypatch.setenv("SPICE_SINK_IDS", "4096")  # ~13 ids per step: the id ring wraps several times
    monkeypatch.setenv("SPICE_SINK_STEPS", "128")
    net, _ = brunel(**kw)
    net.raster_enable(True)
    got_c, got_i = [], []
    issued = read = 0
    rng = np.random.default_rng(5)
    while read < 600:
        while issued < 600 and issued - read < 45:
            k = int(min(rng.integers(1, 31), 600 - issued))
            net.step(k)
            issued += k
        k = int(min(rng.integers(1, 40), issued - read))
        c, i = net.raster_read(k)
        assert c.shape[0] == k
        got_c.append(c)
        got_i.append(i)
        read += k
    assert np.array_equal(np.concatenate(got_c), counts0)
    assert np.array_equal(np.concatenate(got_i), ids0)
    # an unread log that outgrows the ring is reported, not overwritten silently
    net.step(300)
    with pytest.raises(sp.SpiceError):
        net.raster_read()


def test_sink_large_population_sorted(sp):
    """A population wider than one bitmap chunk (> 1,015,808 neurons): lists stay ascending and equal spikes(0)."""
    net = sp.snn(1e-4, 15e-4, (7,))
    P = net.add_population("brunel.poisson", 2_300_000)
    net.raster_enable(True)
    per_step = []
    for _ in range(5):
        net.step(1)
        per_step.append(np.array(P.spikes(0)))
    counts, ids = net.raster_read()
    assert [int(c) for c in counts[:, 0]] == [len(x) for x in per_step]
    assert np.array_equal(ids, np.concatenate(per_step))
    for x in per_step:
        assert len(x) > 3000 and np.all(np.diff(x) > 0)


def test_host_fed_population_replays_poisson_input(sp):
    """A host-fed population (per-population update(), concepts.h:46-57) that replays the spike train the
    Poisson population of a Brunel network emitted drives E and I to exactly the same rasters and state."""
    from spice2_b200 import fixed_probability

    N, p, dt, delay = 4000, 0.1, 1e-4, 15e-4
    w_exc, w_inh = np.float32(2.0 / 400), np.float32(-10.0 / 400)

    def build(host_feed=None):
        net = sp.snn(dt, delay, (1337,))
        P = net.add_host_population(N // 2, host_feed) if host_feed else net.add_population("brunel.poisson", N // 2)
        E = net.add_population("brunel.lif", N * 4 // 10)
        I = net.add_population("brunel.lif", N // 10)
        for (s, d, w) in ((P, E, w_exc), (P, I, w_exc), (E, E, w_exc), (E, I, w_exc), (I, E, w_inh), (I, I, w_inh)):
            net.connect("brunel.fixed_weight", s, d, fixed_probability(p), delay, weight=w)
        return net, (P, E, I)

    steps = 200
    net, pops = build()
    counts, ids = gpu_raster(net, steps, 37)
    state_e = pops[1].get_neurons()
    # P's spike train, step by step
    train, at = [], 0
    for s in range(steps):
        train.append(ids[at: at + counts[s, 0]].copy())
        at += int(counts[s].sum())
    feed = iter(train)
    net2, pops2 = build(lambda _dt: next(feed))
    counts2, ids2 = gpu_raster(net2, steps, 23)
    assert np.array_equal(counts, counts2) and np.array_equal(ids, ids2)
    assert np.array_equal(state_e, pops2[1].get_neurons())
    assert counts[:, 1].sum() > 50  # E does fire in this window


_SPLIT_SCRIPT = """
import sys
import numpy as np
sys.path.insert(0, {root!r})
from spice2_b200.samples import brunel, vogels
out = {{}}
for name, make, steps in (("brunel", brunel, 300), ("vogels", vogels, 1500)):
    net, pops = make()
    net.raster_enable(True)
    net.step(steps)
    counts, ids = net.raster_read()
    out[name + "_counts"], out[name + "_ids"] = counts, ids
    out[name + "_state"] = np.frombuffer(pops[1].get_neurons().tobytes(), dtype=np.uint8)
    out[name + "_events"] = np.array([net.stats()["synaptic_events"]])
# dense vogels: every volley is 3200 + 800 spikes in one step (4 rounds per unit) and half of the runs
# are longer than one warp-wide load (256-target tiles at p = 0.5: 128 entries on average)
net, pops = vogels(p=0.5)
net.raster_enable(True)
net.step(200)
counts, ids = net.raster_read()
out["dense_counts"], out["dense_ids"] = counts, ids
out["dense_E"] = np.frombuffer(pops[0].get_neurons().tobytes(), dtype=np.uint8)
out["dense_I"] = np.frombuffer(pops[1].get_neurons().tobytes(), dtype=np.uint8)
net.step(8)
out["dense_events"] = np.array([net.stats()["synaptic_events"]])
np.savez({dest!r}, **out)
"""


@pytest.mark.parametrize("split", [8, 16])
def test_delivery_cta_shapes(sp, orc, golden, tmp_path, split):
    """Both CTA shapes of the delivery kernel (deliver.cu: 2 x 8 warps per SM, or 16 warps on a unit for
    windows with few, long units) give the golden rasters.  The vogels run starts with one volley of all
    4000 neurons; the dense variant (p = 0.5) has 3200-spike volleys and runs of 128 groups, four times a
    landing slot (the count_stage tail path)."""
    import os
    import subprocess
    import sys
    from pathlib import Path

    root = str(Path(__file__).resolve().parent.parent)
    dest = str(tmp_path / f"split{split}.npz")
    env = dict(os.environ, SPICE_DELIVER_WARPS=str(split))
    subprocess.run([sys.executable, "-c", _SPLIT_SCRIPT.format(root=root, dest=dest)], check=True, env=env, timeout=600)
    got = np.load(dest)
    for name, key in (("brunel", "brunel_300_strict"), ("vogels", "vogels_1500_strict")):
        g = golden["samples"][key]
        assert [int(x) for x in got[name + "_counts"].sum(0)] == g["totals"]
        assert hx(orc.fnv(got[name + "_counts"])) == g["fnv_counts"]
        assert hx(orc.fnv(got[name + "_ids"])) == g["fnv_ids"]
    assert int(got["vogels_events"][0]) > 0
    onet, opops = vogels_oracle(orc, p=0.5)
    rows, ocounts = run_raster(onet, opops, 200)
    assert ocounts.max() == 3200  # whole-population volleys
    assert np.array_equal(got["dense_counts"], ocounts)
    assert np.array_equal(got["dense_ids"], flatten_raster(rows))
    for key, pop in (("dense_E", opops[0]), ("dense_I", opops[1])):
        assert got[key].tobytes() == onet.neurons(pop).tobytes()
    for _ in range(8 + 7):  # see test_synaptic_event_count
        onet.step()
    assert int(got["dense_events"][0]) == onet.events()


@pytest.mark.parametrize("plastic", [False, True])
def test_get_set_neurons_round_trip_changes_nothing(sp, plastic):
    """get_neurons() hands out the state with the pending deliveries of the next step applied (the reference's live
    state, neuron_population.h:142-145); writing the same values back must not apply them a second time."""
    from spice2_b200.samples import brunel

    kw = dict(N=3000, p=0.1, w_exc=np.float32(2.0 / 300), w_inh=np.float32(-10.0 / 300), seed=(9,), plastic=plastic)
    rasters = []
    for round_trip in (False, True):
        net, pops = brunel(**kw)
        net.raster_enable(True)
        net.step(130)
        if round_trip:
            for pi in (1, 2):
                pops[pi].set_neurons(pops[pi].get_neurons())
        net.step(130)
        counts, ids = net.raster_read(260)
        rasters.append((counts, ids, pops[1].get_neurons(), pops[2].get_neurons()))
        net.close()
    for a, b in zip(*rasters):
        assert np.array_equal(a, b)
    assert rasters[0][0][131:, 1:].sum() > 0  # the populations that were read and written did fire afterwards
