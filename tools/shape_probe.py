#!/usr/bin/env python
"""Delivery-kernel efficiency as a function of the network's SHAPE on one GPU: the per-rank problem of an
N-GPU run has more source rows, fewer target tiles per row and longer spike lists per step than the 1-GPU
problem with the same number of synapses.  usage: shape_probe.py P E I [p] [steps]"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import spice2_b200 as sp  # noqa: E402
from spice2_b200 import fixed_probability  # noqa: E402

nP, nE, nI = (int(float(x)) for x in sys.argv[1:4])
p = float(sys.argv[4]) if len(sys.argv) > 4 else 0.02
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 900
dt, delay = 1e-4, 15e-4
net = sp.snn(dt, delay, (1337,))
P = net.add_population("brunel.poisson", nP)
E = net.add_population("brunel.lif", nE)
I = net.add_population("brunel.lif", nI)
N = 2 * nP  # the Brunel proportions' nominal size: weights as in the benchmark
w_exc, w_inh = np.float32(0.2 / (p * N)), np.float32(-1.0 / (p * N))
for (s, d, w) in ((P, E, w_exc), (P, I, w_exc), (E, E, w_exc), (E, I, w_exc), (I, E, w_inh), (I, I, w_inh)):
    net.connect("brunel.fixed_weight", s, d, fixed_probability(p), delay, weight=w)
net.step(300)
net.sync()
s0 = net.stats()
net.profile_enable(True)
net.profile_read()
net.step(steps)
net.sync()
prof = net.profile_read()
s1 = net.stats()
ev = s1["synaptic_events"] - s0["synaptic_events"]
spk = s1["spikes_delivered"] - s0["spikes_delivered"]
alg = 4.0 * ev + 20.0 * spk
print(json.dumps({"P": nP, "E": nE, "I": nI, "p": p, "synapses": sum(net.connection_edges(c) for c in range(6)),
                  "events_per_step": ev / steps, "spikes_per_step": spk / steps, "deliver_us_per_window": prof["deliver_ms"] / prof["windows"] * 1e3,
                  "update_us_per_window": prof["update_ms"] / prof["windows"] * 1e3, "deliver_GBps": alg / (prof["deliver_ms"] * 1e-3) / 1e9}))
