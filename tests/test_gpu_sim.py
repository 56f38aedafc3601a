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
