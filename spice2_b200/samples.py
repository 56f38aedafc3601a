"""The reference's sample networks built through the host mirror (same construction order, hence
the same seed bookkeeping, as samples/brunel.cpp:78-103 and samples/vogels.cpp:62-76)."""
from __future__ import annotations

import numpy as np

from . import fixed_probability, snn


def brunel(N=20000, p=0.1, w_exc=None, w_inh=None, dt=1e-4, delay=15e-4, seed=(1337,), plastic=False, fast_topology=False, partition=None, **ctx):
    """P (poisson, N/2), E (lif, 4N/10), I (lif, N/10); P->E, P->I, E->E, E->I, I->E, I->I.
    plastic=True: E->E carries the STDP synapse of samples/brunel+.cpp:59-99,114.
    fast_topology=True: the connections are drawn by the counter-based generator (not the reference's matrices).
    partition: size -> the ranks' target ranges (world + 1 bounds) instead of equal widths."""
    w_exc = np.float32(2.0 / N) if w_exc is None else np.float32(w_exc)
    w_inh = np.float32(-10.0 / N) if w_inh is None else np.float32(w_inh)
    net = snn(dt, delay, seed, **ctx)
    bounds = (lambda n: None) if partition is None else partition
    P = net.add_population("brunel.poisson", N // 2, bounds=bounds(N // 2))
    E = net.add_population("brunel.lif", N * 4 // 10, bounds=bounds(N * 4 // 10))
    I = net.add_population("brunel.lif", N // 10, bounds=bounds(N // 10))
    for (s, d, w) in ((P, E, w_exc), (P, I, w_exc), (E, E, w_exc), (E, I, w_exc), (I, E, w_inh), (I, I, w_inh)):
        if plastic and s is E and d is E:
            net.connect("brunel+.plastic", s, d, fixed_probability(p, fast_topology), delay)
        else:
            net.connect("brunel.fixed_weight", s, d, fixed_probability(p, fast_topology), delay, weight=w)
    return net, (P, E, I)


def brunel_scaled(N, p=0.02, **kw):
    """Brunel with weights scaled with the in-degree so the 20k-neuron dynamics are
    kept: w_exc = 0.2/(p N), w_inh = -1.0/(p N) (equal to 2/N, -10/N at p = 0.1)."""
    return brunel(N=N, p=p, w_exc=np.float32(0.2 / (p * N)), w_inh=np.float32(-1.0 / (p * N)), **kw)


def vogels(N=4000, p=0.02, w_exc=None, w_inh=None, dt=1e-4, delay=8e-4, seed=(1337,), **ctx):
    """E (lif, 8N/10), I (lif, 2N/10); E->E, E->I excitatory, I->E, I->I inhibitory."""
    w_exc = np.float32(6.4e6 / (N * N)) if w_exc is None else np.float32(w_exc)
    w_inh = np.float32(8.16e7 / (N * N)) if w_inh is None else np.float32(w_inh)
    net = snn(dt, delay, seed, **ctx)
    E = net.add_population("vogels.lif", N * 8 // 10)
    I = net.add_population("vogels.lif", N * 2 // 10)
    net.connect("vogels.excitatory", E, E, fixed_probability(p), delay, weight=w_exc)
    net.connect("vogels.excitatory", E, I, fixed_probability(p), delay, weight=w_exc)
    net.connect("vogels.inhibitory", I, E, fixed_probability(p), delay, weight=w_inh)
    net.connect("vogels.inhibitory", I, I, fixed_probability(p), delay, weight=w_inh)
    return net, (E, I)
