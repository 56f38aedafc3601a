#!/usr/bin/env python
"""The delivery problem ONE RANK of a G-GPU Brunel run has, reproduced on one GPU: sources are whole
populations (host-fed populations emitting random spikes at the benchmark's rates), targets are the rank's
share (LIF populations of size/G that receive with zero weight, so they never fire).  Prints the delivery
kernel's algorithmic GB/s for that shape.  usage: rank_shape_probe.py G [steps] [burst]
burst (0..1) modulates every population's rate by 1 + burst * sin(2 pi step / 10): the population-wide
oscillation a recurrent Brunel network shows, which makes the units of some steps several times longer than
those of others.  SPICE_DELIVER_SPLIT=0/1 forces whole-unit / single-round work items (deliver.cu)."""
import os
import json
import math
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import spice2_b200 as sp  # noqa: E402
from spice2_b200 import fixed_probability  # noqa: E402

G = int(sys.argv[1])
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
burst = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
n = int(round(2.0e6 * math.sqrt(G / 8.0) / (10 * G))) * 10 * G
nP, nE, nI = n // 2, n * 4 // 10, n // 10
p, dt, delay = 0.02, 1e-4, 15e-4
rates = {"P": 20.0, "E": 36.0, "I": 36.0}
rng = np.random.default_rng(1)


def feeder(size, rate):
    t = [0]

    def f(_dt):
        k = rng.poisson(size * rate * dt * (1.0 + burst * math.sin(2.0 * math.pi * t[0] / 10.0)))
        t[0] += 1
        return np.sort(rng.choice(size, size=min(k, size), replace=False)).astype(np.int32)
    return f


net = sp.snn(dt, delay, (1337,))
P = net.add_host_population(nP, feeder(nP, rates["P"]))
Es = net.add_host_population(nE, feeder(nE, rates["E"]))
Is = net.add_host_population(nI, feeder(nI, rates["I"]))
Ed = net.add_population("brunel.lif", nE // G)
Id = net.add_population("brunel.lif", nI // G)
zero = np.float32(0.0)
for (s, d) in ((P, Ed), (P, Id), (Es, Ed), (Es, Id), (Is, Ed), (Is, Id)):
    net.connect("brunel.fixed_weight", s, d, fixed_probability(p), delay, weight=zero)
net.step(45)
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
print(json.dumps({"ranks_emulated": G, "burst": burst, "split": os.environ.get("SPICE_DELIVER_SPLIT", "auto"),
                  "sources": [nP, nE, nI], "targets": [nE // G, nI // G],
                  "synapses": sum(net.connection_edges(c) for c in range(6)), "events_per_step": ev / steps,
                  "spikes_per_step": spk / steps, "deliver_us_per_window": prof["deliver_ms"] / prof["windows"] * 1e3,
                  "deliver_GBps": alg / (prof["deliver_ms"] * 1e-3) / 1e9}))
