"""ctypes bindings for the parity checkers (TEST INFRASTRUCTURE).

* ``Oracle``  -> oracle/liboracle.so, the C restatement (oracle/spice_oracle.c);
* ``RefShim`` -> oracle/_ref/libspice_ref_{fast,strict}.so, the UNMODIFIED reference compiled
  from /root/reference by oracle/Makefile (present in the build container and shipped to the
  GPU box as a prebuilt file; absent elsewhere -> tests that need it skip).

Nothing under spice2_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"

POISSON, LIF_BRUNEL, LIF_VOGELS = 0, 1, 2
FIXED_WEIGHT_V, WEIGHT_GEX, WEIGHT_GIN, PLASTIC_BRUNEL = 0, 1, 2, 3
STRICT, REFBUILD = 0, 1

LIF_BRUNEL_DT = np.dtype([("V", np.float32), ("Twait", np.int32)])
LIF_VOGELS_DT = np.dtype([("V", np.float32), ("Gex", np.float32), ("Gin", np.float32), ("Twait", np.int32)])
SYN_PLASTIC_DT = np.dtype([("W", np.float32), ("Zpre", np.float32), ("Zpost", np.float32)])


class U128(C.Structure):
    _fields_ = [("lo", C.c_uint64), ("hi", C.c_uint64)]

    def tup(self):
        return (int(self.lo), int(self.hi))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def build_oracle():
    subprocess.run(["make", "-s", "-C", str(ORACLE_DIR), "oracle"], check=True)


class Oracle:
    def __init__(self):
        so = ORACLE_DIR / "liboracle.so"
        if not so.exists():
            build_oracle()
        L = self.L = C.CDLL(str(so))
        L.orc_seed_seq.restype = U128
        L.orc_seed_seq.argtypes = [C.c_void_p, C.c_int]
        L.orc_seed_next.restype = U128
        L.orc_seed_next.argtypes = [U128]
        L.orc_seed_stream.restype = U128
        L.orc_seed_stream.argtypes = [U128, C.c_uint64]
        L.orc_xoroshiro.argtypes = [U128, C.c_int64, C.c_void_p]
        L.orc_xoroshiro_state_at.argtypes = [U128, C.c_int64, C.c_void_p]
        L.orc_xoroshiro_jump.argtypes = [U128, C.c_uint64, C.c_void_p]
        L.orc_fixed_probability_rows_from.restype = C.c_int64
        L.orc_fixed_probability_rows_from.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_double, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                                      C.c_void_p, C.c_int64]
        L.orc_kahan_dt.argtypes = [C.c_float, C.c_int64, C.c_void_p]
        L.orc_fnv1a64.restype = C.c_uint64
        L.orc_fnv1a64.argtypes = [C.c_void_p, C.c_int64]
        L.orc_fixed_probability_max_degree.restype = C.c_int64
        L.orc_fixed_probability_max_degree.argtypes = [C.c_int64, C.c_double]
        L.orc_fixed_probability_size.restype = C.c_int64
        L.orc_fixed_probability_size.argtypes = [C.c_int64, C.c_int64, C.c_double]
        L.orc_fixed_probability_generate.restype = C.c_int64
        L.orc_fixed_probability_generate.argtypes = [C.c_int64, C.c_int64, C.c_double, U128, C.c_void_p,
                                                     C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_net_create.restype = C.c_void_p
        L.orc_net_create.argtypes = [C.c_float, C.c_float, C.c_void_p, C.c_int, C.c_int]
        L.orc_net_destroy.argtypes = [C.c_void_p]
        L.orc_net_set_shard.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_add_population.argtypes = [C.c_void_p, C.c_int, C.c_int64]
        L.orc_connect.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_float, C.c_int, C.c_float]
        for f in ("orc_step", "orc_step_update", "orc_step_deliver"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_step_set_spikes.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]
        for f in ("orc_population_size", "orc_population_lo", "orc_population_hi"):
            getattr(L, f).restype = C.c_int64
            getattr(L, f).argtypes = [C.c_void_p, C.c_int]
        L.orc_spikes.restype = C.c_int64
        L.orc_spikes.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.POINTER(C.c_void_p)]
        L.orc_local_spikes.restype = C.c_int64
        L.orc_local_spikes.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        L.orc_neurons.restype = C.c_void_p
        L.orc_neurons.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]
        L.orc_synaptic_events.restype = C.c_int64
        L.orc_synaptic_events.argtypes = [C.c_void_p]
        L.orc_connection_edges.restype = C.c_int64
        L.orc_connection_edges.argtypes = [C.c_void_p, C.c_int]
        L.orc_connection_offsets.restype = C.c_void_p
        L.orc_connection_offsets.argtypes = [C.c_void_p, C.c_int]
        L.orc_connection_neighbors.restype = C.c_void_p
        L.orc_connection_neighbors.argtypes = [C.c_void_p, C.c_int]
        L.orc_connection_synapses.restype = C.c_void_p
        L.orc_connection_synapses.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]

    # -- seeds / rng ----------------------------------------------------------------------
    def seed_seq(self, il, increments=0) -> U128:
        a = np.asarray(il, np.uint32)
        s = self.L.orc_seed_seq(_ptr(a), len(a))
        for _ in range(increments):
            s = self.L.orc_seed_next(s)
        return s

    def seed_next(self, s):
        return self.L.orc_seed_next(s)

    def xoroshiro(self, seed: U128, count):
        out = np.zeros(count, np.uint64)
        self.L.orc_xoroshiro(seed, count, _ptr(out))
        return out

    def state_at(self, seed: U128, k):
        out = np.zeros(2, np.uint64)
        self.L.orc_xoroshiro_state_at(seed, k, _ptr(out))
        return int(out[0]), int(out[1])

    def jump(self, seed: U128, k):
        """Engine state after k draws by GF(2) matrix powers (no walk)."""
        out = np.zeros(2, np.uint64)
        self.L.orc_xoroshiro_jump(seed, C.c_uint64(k), _ptr(out))
        return int(out[0]), int(out[1])

    def kahan_dt(self, dt, steps):
        out = np.zeros(steps, np.float32)
        self.L.orc_kahan_dt(dt, steps, _ptr(out))
        return out

    def fnv(self, a: np.ndarray) -> int:
        a = np.ascontiguousarray(a)
        return int(self.L.orc_fnv1a64(_ptr(a), a.nbytes))

    # -- fixed_probability ------------------------------------------------------------------
    def max_degree(self, dst, p):
        return int(self.L.orc_fixed_probability_max_degree(dst, p))

    def fixed_probability(self, src, dst, p, seed: U128, want_neighbors=True, want_row_hash=False):
        cap = int(self.L.orc_fixed_probability_size(src, dst, p))
        offsets = np.zeros(src + 1, np.int64)
        nb = np.zeros(max(cap, 1), np.int32) if want_neighbors else None
        rh = np.zeros(max(src, 1), np.uint64) if want_row_hash else None
        draws = C.c_int64()
        e = int(self.L.orc_fixed_probability_generate(src, dst, p, seed, _ptr(offsets), _ptr(nb), _ptr(rh),
                                                      C.byref(draws)))
        return dict(edges=e, offsets=offsets, neighbors=None if nb is None else nb[:e], row_hash=rh,
                    draws=int(draws.value), capacity=cap)

    def fixed_probability_rows(self, state, rows, dst, p, col_lo=0, col_hi=None, want_neighbors=True):
        """`rows` consecutive rows from the engine state at the first row's first draw; targets in [col_lo, col_hi) kept
        as local columns -> dict(kept_degree, full_degree, neighbors)."""
        col_hi = dst if col_hi is None else col_hi
        st = np.asarray(state, np.uint64)
        kd, fd = np.zeros(rows, np.int64), np.zeros(rows, np.int64)
        cap = int(rows * min(self.max_degree(dst, p), col_hi - col_lo)) if want_neighbors else 0
        nb = np.zeros(max(cap, 1), np.int32) if want_neighbors else None
        self.L.orc_fixed_probability_rows_from.restype = C.c_int64
        n = int(self.L.orc_fixed_probability_rows_from(_ptr(st), C.c_int64(rows), C.c_int64(dst), C.c_double(p), C.c_int64(col_lo),
                                                       C.c_int64(col_hi), _ptr(kd), _ptr(fd), _ptr(nb), C.c_int64(cap)))
        return dict(kept=n, kept_degree=kd, full_degree=fd, neighbors=None if nb is None else nb[:n])

    # -- networks ---------------------------------------------------------------------------
    def net(self, dt, max_delay, seed_il=(1337,), flavour=STRICT, rank=0, world=1):
        return OracleNet(self, dt, max_delay, seed_il, flavour, rank, world)


class OracleNet:
    def __init__(self, orc: Oracle, dt, max_delay, seed_il, flavour, rank, world):
        self.L = orc.L
        il = np.asarray(seed_il, np.uint32)
        self.h = self.L.orc_net_create(dt, max_delay, _ptr(il), len(il), flavour)
        self.L.orc_net_set_shard(self.h, rank, world)
        self.models = []

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_net_destroy(self.h)
            self.h = None

    def add_population(self, model, size):
        self.models.append(model)
        return self.L.orc_add_population(self.h, model, size)

    def connect(self, src, dst, p, delay, syn_model, weight=0.0):
        r = self.L.orc_connect(self.h, src, dst, p, delay, syn_model, weight)
        if r != 0:
            raise ValueError("precondition violated: 1 <= round(delay/dt) <= max_delay")

    def step(self):
        self.L.orc_step(self.h)

    def step_update(self):
        self.L.orc_step_update(self.h)

    def step_set_spikes(self, pop, ids):
        ids = np.ascontiguousarray(ids, np.int32)
        self.L.orc_step_set_spikes(self.h, pop, _ptr(ids), len(ids))

    def step_deliver(self):
        self.L.orc_step_deliver(self.h)

    def spikes(self, pop, age=0):
        p = C.c_void_p()
        n = self.L.orc_spikes(self.h, pop, age, C.byref(p))
        if n < 0:
            raise ValueError("age out of range")
        if n == 0:
            return np.zeros(0, np.int32)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int32)), shape=(n,)).copy()

    def local_spikes(self, pop):
        p = C.c_void_p()
        n = self.L.orc_local_spikes(self.h, pop, C.byref(p))
        if n == 0:
            return np.zeros(0, np.int32)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int32)), shape=(n,)).copy()

    def bounds(self, pop):
        return int(self.L.orc_population_lo(self.h, pop)), int(self.L.orc_population_hi(self.h, pop))

    def neurons(self, pop):
        b = C.c_int64()
        p = self.L.orc_neurons(self.h, pop, C.byref(b))
        lo, hi = self.bounds(pop)
        if not p or b.value == 0:
            return None
        dt = LIF_BRUNEL_DT if self.models[pop] == LIF_BRUNEL else LIF_VOGELS_DT
        raw = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=((hi - lo) * b.value,)).copy()
        return raw.view(dt)

    def events(self):
        return int(self.L.orc_synaptic_events(self.h))

    def connection(self, ci):
        e = int(self.L.orc_connection_edges(self.h, ci))
        return e

    def connection_csr(self, ci, src_size):
        e = int(self.L.orc_connection_edges(self.h, ci))
        po = self.L.orc_connection_offsets(self.h, ci)
        pn = self.L.orc_connection_neighbors(self.h, ci)
        off = np.ctypeslib.as_array(C.cast(po, C.POINTER(C.c_int64)), shape=(src_size + 1,)).copy()
        nb = np.ctypeslib.as_array(C.cast(pn, C.POINTER(C.c_int32)), shape=(max(e, 1),)).copy()[:e]
        return off, nb

    def connection_synapses(self, ci):
        e = int(self.L.orc_connection_edges(self.h, ci))
        b = C.c_int64()
        p = self.L.orc_connection_synapses(self.h, ci, C.byref(b))
        if not p:
            return None
        raw = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(e * b.value,)).copy()
        return raw.view(SYN_PLASTIC_DT)


def brunel_oracle(orc: Oracle, N=20000, p=0.1, w_exc=None, w_inh=None, dt=1e-4, delay=15e-4, seed=(1337,),
                  flavour=STRICT, plastic=False, rank=0, world=1):
    """samples/brunel.cpp:78-103 (plastic=True: samples/brunel+.cpp:102-117)."""
    w_exc = np.float32(2.0 / N) if w_exc is None else np.float32(w_exc)
    w_inh = np.float32(-10.0 / N) if w_inh is None else np.float32(w_inh)
    net = orc.net(np.float32(dt), np.float32(delay), seed, flavour, rank, world)
    P = net.add_population(POISSON, N // 2)
    E = net.add_population(LIF_BRUNEL, N * 4 // 10)
    I = net.add_population(LIF_BRUNEL, N // 10)
    d = np.float32(delay)
    net.connect(P, E, p, d, FIXED_WEIGHT_V, w_exc)
    net.connect(P, I, p, d, FIXED_WEIGHT_V, w_exc)
    if plastic:
        net.connect(E, E, p, d, PLASTIC_BRUNEL)
    else:
        net.connect(E, E, p, d, FIXED_WEIGHT_V, w_exc)
    net.connect(E, I, p, d, FIXED_WEIGHT_V, w_exc)
    net.connect(I, E, p, d, FIXED_WEIGHT_V, w_inh)
    net.connect(I, I, p, d, FIXED_WEIGHT_V, w_inh)
    return net, (P, E, I)


def vogels_oracle(orc: Oracle, N=4000, p=0.02, w_exc=None, w_inh=None, dt=1e-4, delay=8e-4, seed=(1337,),
                  rank=0, world=1):
    """samples/vogels.cpp:62-76."""
    w_exc = np.float32(6.4e6 / (N * N)) if w_exc is None else np.float32(w_exc)
    w_inh = np.float32(8.16e7 / (N * N)) if w_inh is None else np.float32(w_inh)
    net = orc.net(np.float32(dt), np.float32(delay), seed, STRICT, rank, world)
    E = net.add_population(LIF_VOGELS, N * 8 // 10)
    I = net.add_population(LIF_VOGELS, N * 2 // 10)
    d = np.float32(delay)
    net.connect(E, E, p, d, WEIGHT_GEX, w_exc)
    net.connect(E, I, p, d, WEIGHT_GEX, w_exc)
    net.connect(I, E, p, d, WEIGHT_GIN, w_inh)
    net.connect(I, I, p, d, WEIGHT_GIN, w_inh)
    return net, (E, I)


def run_raster(net: OracleNet, pops, steps):
    """-> (ids per (step,pop) list of arrays, counts[steps,npop])"""
    rows = []
    counts = np.zeros((steps, len(pops)), np.int64)
    for s in range(steps):
        net.step()
        r = [net.spikes(p, 0) for p in pops]
        rows.append(r)
        counts[s] = [len(x) for x in r]
    return rows, counts


# ---------------------------------------------------------------------------------------------
class RefShim:
    """The compiled reference (oracle/_ref).  flavour: 'fast' = the reference's own flags,
    'strict' = IEEE evaluation of the same sources."""

    @staticmethod
    def available(flavour="fast"):
        return (ORACLE_DIR / "_ref" / f"libspice_ref_{flavour}.so").exists()

    def __init__(self, flavour="fast"):
        self.L = C.CDLL(str(ORACLE_DIR / "_ref" / f"libspice_ref_{flavour}.so"))
        self.L.ref_fixed_probability_size.restype = C.c_int64
        self.L.ref_fixed_probability_generate.restype = C.c_int64
        self.L.ref_build_flavour.restype = C.c_char_p

    def seed(self, il, increments=0):
        a = np.asarray(il, np.uint32)
        out = (C.c_uint64 * 2)()
        self.L.ref_seed(_ptr(a), C.c_int(len(a)), C.c_int(increments), out)
        return int(out[0]), int(out[1])

    def xoroshiro(self, il, increments, count):
        a = np.asarray(il, np.uint32)
        out = np.zeros(count, np.uint64)
        self.L.ref_xoroshiro(_ptr(a), C.c_int(len(a)), C.c_int(increments), C.c_int64(count), _ptr(out))
        return out

    def canonical_float(self, il, increments, count):
        a = np.asarray(il, np.uint32)
        out = np.zeros(count, np.float32)
        self.L.ref_canonical_float(_ptr(a), C.c_int(len(a)), C.c_int(increments), C.c_int64(count), _ptr(out))
        return out

    def exponential(self, il, increments, scale, count):
        a = np.asarray(il, np.uint32)
        out = np.zeros(count, np.float64)
        self.L.ref_exponential(_ptr(a), C.c_int(len(a)), C.c_int(increments), C.c_double(scale),
                               C.c_int64(count), _ptr(out))
        return out

    def random_sample(self, kind, il, increments, a, b, count):
        """oracle/ref_shim.cpp ref_random_sample: the reference's 32-bit engine, normal / binomial / float exponential
        distributions and seed_seq(std::seed_seq), as doubles."""
        w = np.asarray(il, np.uint32)
        out = np.zeros(count, np.float64)
        self.L.ref_random_sample(C.c_int(kind), _ptr(w), C.c_int(len(w)), C.c_int(increments), C.c_double(a), C.c_double(b),
                                 C.c_int64(count), _ptr(out))
        return out

    def kahan_dt(self, dt, steps):
        out = np.zeros(steps, np.float32)
        self.L.ref_kahan_dt(C.c_float(dt), C.c_int64(steps), _ptr(out))
        return out

    def libm_log(self, x):
        x = np.ascontiguousarray(x, np.float64)
        out = np.zeros_like(x)
        self.L.ref_libm_log(_ptr(x), C.c_int64(len(x)), _ptr(out))
        return out

    def libm_expf(self, x):
        x = np.ascontiguousarray(x, np.float32)
        out = np.zeros_like(x)
        self.L.ref_libm_expf(_ptr(x), C.c_int64(len(x)), _ptr(out))
        return out

    def libm_pow(self, x, y):
        x = np.ascontiguousarray(x, np.float64)
        y = np.ascontiguousarray(y, np.float64)
        out = np.zeros_like(x)
        self.L.ref_libm_pow(_ptr(x), _ptr(y), C.c_int64(len(x)), _ptr(out))
        return out

    def fixed_probability(self, src, dst, p, il=(1337,), increments=0):
        a = np.asarray(il, np.uint32)
        cap = int(self.L.ref_fixed_probability_size(C.c_int64(src), C.c_int64(dst), C.c_double(p)))
        offsets = np.zeros(src + 1, np.int64)
        nb = np.zeros(max(cap, 1), np.int32)
        sec = C.c_double()
        e = int(self.L.ref_fixed_probability_generate(C.c_int64(src), C.c_int64(dst), C.c_double(p), _ptr(a),
                                                      C.c_int(len(a)), C.c_int(increments), _ptr(offsets), _ptr(nb),
                                                      C.byref(sec)))
        return dict(edges=e, offsets=offsets, neighbors=nb[:e], capacity=cap, seconds=sec.value)

    def _run(self, fn, N, p, w_exc, w_inh, dt, delay, seed, steps, npop, nE, nI, state_dt, record=True,
             capacity=None):
        cap = capacity or max(1 << 20, int(steps * N * 0.02))
        ids = np.zeros(cap, np.int32) if record else None
        counts = np.zeros(steps * npop, np.int64) if record else None
        sE = np.zeros(nE, state_dt)
        sI = np.zeros(nI, state_dt)
        b, s = C.c_double(), C.c_double()
        args = [C.c_int64(N), C.c_double(p), C.c_float(w_exc), C.c_float(w_inh), C.c_float(dt), C.c_float(delay),
                C.c_uint32(seed), C.c_int64(steps), _ptr(ids), C.c_int64(cap), _ptr(counts), _ptr(sE), _ptr(sI),
                C.byref(b), C.byref(s)]
        ev = C.c_int64(-1)
        if fn == "ref_brunel_run":
            args.append(C.byref(ev) if record else None)
        r = getattr(self.L, fn)(*args)
        if r != 0:
            raise RuntimeError("raster capacity exceeded")
        out = dict(build_seconds=b.value, sim_seconds=s.value, state_E=sE, state_I=sI, synaptic_events=ev.value)
        if record:
            counts = counts.reshape(steps, npop)
            out["counts"] = counts
            out["ids"] = ids[: int(counts.sum())].copy()
        return out

    def brunel(self, N=20000, p=0.1, w_exc=None, w_inh=None, dt=1e-4, delay=15e-4, seed=1337, steps=300,
               plastic=False, record=True):
        w_exc = np.float32(2.0 / N) if w_exc is None else np.float32(w_exc)
        w_inh = np.float32(-10.0 / N) if w_inh is None else np.float32(w_inh)
        fn = "ref_brunel_plus_run" if plastic else "ref_brunel_run"
        return self._run(fn, N, p, w_exc, w_inh, np.float32(dt), np.float32(delay), seed, steps, 3, N * 4 // 10,
                         N // 10, LIF_BRUNEL_DT, record)

    def brunel_open(self, N, p, w_exc, w_inh, dt=1e-4, delay=15e-4, seed=1337):
        """Incremental Brunel run of the compiled reference (bench.py's reference arm): -> BrunelRun."""
        return BrunelRun(self.L, N, p, w_exc, w_inh, dt, delay, seed)

    def vogels(self, N=4000, p=0.02, w_exc=None, w_inh=None, dt=1e-4, delay=8e-4, seed=1337, steps=1500,
               record=True):
        w_exc = np.float32(6.4e6 / (N * N)) if w_exc is None else np.float32(w_exc)
        w_inh = np.float32(8.16e7 / (N * N)) if w_inh is None else np.float32(w_inh)
        return self._run("ref_vogels_run", N, p, w_exc, w_inh, np.float32(dt), np.float32(delay), seed, steps, 2,
                         N * 8 // 10, N * 2 // 10, LIF_VOGELS_DT, record)


class BrunelRun:
    """ref_brunel_open / ref_brunel_advance / ref_brunel_close (oracle/ref_shim.cpp)."""

    def __init__(self, L, N, p, w_exc, w_inh, dt, delay, seed):
        self.L = L
        L.ref_brunel_open.restype = C.c_void_p
        b = C.c_double()
        self.h = C.c_void_p(L.ref_brunel_open(C.c_int64(N), C.c_double(p), C.c_float(w_exc), C.c_float(w_inh), C.c_float(dt),
                                              C.c_float(delay), C.c_uint32(seed), C.byref(b)))
        self.build_seconds = b.value

    def advance(self, steps):
        """-> (seconds inside snn::step(), Syn::deliver invocations, spikes emitted) for `steps` time steps"""
        s, ev, sp = C.c_double(), C.c_int64(), C.c_int64()
        self.L.ref_brunel_advance(self.h, C.c_int64(steps), C.byref(s), C.byref(ev), C.byref(sp))
        return s.value, ev.value, sp.value

    def close(self):
        if self.h:
            self.L.ref_brunel_close(self.h)
            self.h = None

    def __del__(self):
        self.close()


def flatten_raster(rows):
    """rows[step][pop] -> flat ids in (step, pop) order, as RefShim returns them."""
    parts = [x for r in rows for x in r]
    return np.concatenate(parts) if parts else np.zeros(0, np.int32)
