"""Generate tests/golden/*.json|*.npz from the COMPILED REFERENCE (oracle/_ref, built by
`make -C oracle ref` from /root/reference).  Run in the build container only:

    python tests/golden/make_golden.py

The fixtures are small and committed; /root/reference and oracle/_ref are not needed to use them.
"""
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
from oracle_lib import Oracle, RefShim  # noqa: E402

orc = Oracle()  # only for its FNV helper
fast, strict = RefShim("fast"), RefShim("strict")
hx = lambda v: f"{int(v):016x}"
G = {"reference": "denniskb/spice2 @ f5e57eb", "flavours": {"fast": fast.L.ref_build_flavour().decode(),
                                                            "strict": strict.L.ref_build_flavour().decode()}}

# ---- seeds / rng / dt ----------------------------------------------------------------------
G["seed_1337"] = {str(k): [hx(x) for x in fast.seed([1337], k)] for k in (0, 1, 2, 8, 100)}
G["seed_il"] = {",".join(map(str, il)): [hx(x) for x in fast.seed(il)]
                for il in ([1], [1, 2], [1, 2, 3], [1, 2, 3, 4], [1, 2, 3, 4, 5], [0xFFFFFFFF, 7])}
G["xoroshiro_1337_first8"] = [hx(x) for x in fast.xoroshiro([1337], 0, 8)]
G["xoroshiro_1337_inc3_fnv_1e5"] = hx(orc.fnv(fast.xoroshiro([1337], 3, 100000)))
dt = fast.kahan_dt(np.float32(1e-4), 30000)
G["kahan_dt_1e-4"] = {"first8": [hx(x) for x in dt[:8].view(np.uint32)], "i299": hx(dt[299:300].view(np.uint32)[0]),
                      "fnv_30000": hx(orc.fnv(dt)), "distinct_300": int(len(np.unique(dt[:300])))}
G["canonical_float_1337_inc8_fnv_1e5"] = hx(orc.fnv(fast.canonical_float([1337], 8, 100000)))
G["exponential_scale9_1337_fnv_1e5"] = hx(orc.fnv(fast.exponential([1337], 0, 9.0, 100000)))

# ---- fixed_probability -----------------------------------------------------------------------
cases = [(1000, 1000, 0.1, 0), (10000, 10000, 0.1, 0), (3000, 777, 0.02, 2), (500, 20000, 0.5, 1),
         (100, 100, 1.0, 0), (64, 5000, 0.001, 0), (7, 3, 0.3, 5), (2000, 2000, 0.9, 0), (1, 1, 0.5, 0),
         (5, 100000, 0.02, 0), (20000, 160, 0.02, 3), (10000, 8000, 0.1, 1), (100000, 100000, 0.1, 0)]
fp = []
for (s, d, p, inc) in cases:
    r = fast.fixed_probability(s, d, p, [1337], inc)
    deg = np.diff(r["offsets"])
    fp.append(dict(src=s, dst=d, p=p, increments=inc, capacity=r["capacity"], edges=r["edges"],
                   deg_min=int(deg.min()), deg_max=int(deg.max()), row0=[int(x) for x in r["neighbors"][:8]],
                   fnv_offsets=hx(orc.fnv(r["offsets"])), fnv_neighbors=hx(orc.fnv(r["neighbors"]))))
    print("fixed_probability", s, d, p, inc, r["edges"], "%.2fs" % r["seconds"], flush=True)
    del r
G["fixed_probability"] = fp

# ---- samples -----------------------------------------------------------------------------------
def summarize(r, steps):
    c = r["counts"]
    return dict(steps=steps, totals=[int(x) for x in c.sum(0)], fnv_counts=hx(orc.fnv(c)), fnv_ids=hx(orc.fnv(r["ids"])),
                fnv_state_E=hx(orc.fnv(r["state_E"])), fnv_state_I=hx(orc.fnv(r["state_I"])),
                first_rows=[[int(x) for x in r["ids"][:int(c[0].sum())][:6]]])

S = {}
S["brunel_300_fast"] = summarize(fast.brunel(steps=300), 300)
S["brunel_300_strict"] = summarize(strict.brunel(steps=300), 300)
S["brunel_3000_fast"] = summarize(fast.brunel(steps=3000), 3000)
S["brunel_3000_strict"] = summarize(strict.brunel(steps=3000), 3000)
S["brunel_40k_300_fast"] = summarize(fast.brunel(N=40000, steps=300), 300)
S["brunel_40k_300_strict"] = summarize(strict.brunel(N=40000, steps=300), 300)
S["vogels_1500_fast"] = summarize(fast.vogels(steps=1500), 1500)
S["vogels_1500_strict"] = summarize(strict.vogels(steps=1500), 1500)
S["brunel_plus_300_fast"] = summarize(fast.brunel(steps=300, plastic=True), 300)
S["brunel_plus_300_strict"] = summarize(strict.brunel(steps=300, plastic=True), 300)
# a small odd-sized net with a shorter delay and another seed, to exercise ragged sizes
S["brunel_small_strict"] = summarize(strict.brunel(N=3010, p=0.07, w_exc=np.float32(2.0 / 300), w_inh=np.float32(-10.0 / 300),
                                                   delay=7e-4, seed=42, steps=400), 400)
for k, v in S.items():
    print(k, v["totals"], flush=True)
G["samples"] = S
# md5 of the reference sample programs' stdout (oracle/_ref/{brunel,brunel+,vogels})
import ctypes as C
import hashlib, subprocess
G["sample_stdout_md5"] = {n: hashlib.md5(subprocess.run([str(HERE.parent.parent / "oracle" / "_ref" / n)], capture_output=True,
                                                         check=True).stdout).hexdigest() for n in ("brunel", "brunel+", "vogels", "ping_pong", "external_input")}
# samples/sssp.cpp run by the compiled reference (DeliverFromTo synapses): distances of the 7 vertices
_d = np.zeros(7, np.int64)
C.CDLL(str(HERE.parent.parent / "oracle" / "_ref" / "libspice_ref_strict.so")).ref_sssp_distances(_d.ctypes.data_as(C.c_void_p))
G["sssp_distances"] = [int(x) for x in _d]
(HERE / "golden.json").write_text(json.dumps(G, indent=1))

# ---- libm pins (glibc 2.39 log / expf / pow at the reference's call sites) ---------------------
rng = np.random.default_rng(1234)
r64 = rng.integers(0, 2**64, 1 << 15, dtype=np.uint64)
u = ((r64 >> np.uint64(11)) + np.uint64(1)).astype(np.float64) * 2.0**-53  # generate_canonical<double,true>
edge = np.array([2.0**-53, 2.0**-52, 1.0, 1.0 - 2.0**-53, 0.5, 0.9375, 1.0 - 2.0**-4, 0.93750000000000011, 0.25,
                 1.0 - 2.0**-20, 0.9999, 0.94, 0.95, 0.99], np.float64)
near1 = 1.0 - rng.integers(1, 2**49, 4096).astype(np.float64) * 2.0**-53   # (1-2^-4, 1): glibc's near-1 branch
x = np.concatenate([u, edge, near1])
xf = np.concatenate([-rng.random(4096).astype(np.float32) * np.float32(40), np.float32([0, -1e-8, -87.0, -100.0, -1.0])])
base = np.float64(np.float32(1) - np.float32(1e-4) * (np.float32(1) / np.float32(0.02)))
pn = np.arange(0, 65, dtype=np.float64)
np.savez_compressed(HERE / "libm_pins.npz", log_x=x, log_y=fast.libm_log(x), expf_x=xf, expf_y=fast.libm_expf(xf),
                    pow_x=np.full_like(pn, base), pow_n=pn, pow_y=fast.libm_pow(np.full_like(pn, base), pn))
print("wrote golden.json, libm_pins.npz")
