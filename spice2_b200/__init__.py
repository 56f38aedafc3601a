"""spice2_b200 — Python host mirror of the reference's `snn` API over the B200 C ABI.

The product is the CUDA library (spice2_b200/csrc -> libspice_b200.so, C ABI in
include/spice_b200.h) and the C++20 facade (csrc/include/spice/snn.h).  This module is the
Python-side mirror used by tests/ and bench.py: same vocabulary as the reference
(`snn(dt, max_delay, seed)`, `add_population`, `connect(..., fixed_probability(p), delay, ...)`,
`step()`, `population.spikes(age)`; reference: spice/include/spice/snn.h:16-75), every call
going straight through ctypes into the C ABI.  There is no CPU fallback: without the CUDA
library or a CUDA device every computing call raises.
"""
from __future__ import annotations

import ctypes as C
import struct
from pathlib import Path

import numpy as np

from .build import build as _build

__all__ = ["snn", "fixed_probability", "adj_list", "SpiceError", "lib", "balance_ranges", "generate_fixed_probability", "generate_adj_list", "Adjacency", "seed_seq", "fnv1a64",
           "MODE_DETERMINISTIC", "MODE_FAST"]

MODE_DETERMINISTIC, MODE_FAST = 0, 1
_ERR = {1: "precondition", 2: "cuda", 3: "unsupported", 4: "internal", 5: "no device"}


class SpiceError(RuntimeError):
    """Non-zero status from the C ABI.  Violated preconditions carry the reference's message form
    'Assertion failed (file:line): cond' (spice/src/util/assert.cpp:8-15)."""

    def __init__(self, code, msg):
        super().__init__(f"[{_ERR.get(code, code)}] {msg}")
        self.code = code


_lib = None


def lib() -> C.CDLL:
    """The CUDA library; built in-tree on first use.  Raises if it cannot be had — never falls
    back to a CPU path."""
    global _lib
    if _lib is None:
        # Several rank contexts in ONE process (tests): a kernel loaded lazily at its first launch can synchronise the device,
        # which never returns while another rank's wait kernel spins for work this thread has not enqueued yet.  Load every
        # kernel when the module is loaded instead (no effect once CUDA is initialised; one process per GPU never needs it).
        import os

        os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
        path = _build()
        L = C.CDLL(str(path))
        vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
        sig = {
            "spice_ctx_create": (i32, [C.POINTER(vp), i32, C.c_float, C.c_float, vp, i32, i32, i32, i32]),
            "spice_ctx_create_seeded": (i32, [C.POINTER(vp), i32, C.c_float, C.c_float, C.c_uint64, C.c_uint64, i32, i32, i32]),
            "spice_ctx_destroy": (i32, [vp]),
            "spice_last_error": (C.c_char_p, [vp]),
            "spice_ctx_set_stream": (i32, [vp, vp]),
            "spice_ctx_seed": (i32, [vp, C.POINTER(C.c_uint64)]),
            "spice_add_population": (i32, [vp, vp, i64, vp, C.POINTER(i32)]),
            "spice_add_host_population": (i32, [vp, i64, vp, vp, C.POINTER(i32)]),
            "spice_population_size": (i64, [vp, i32]),
            "spice_set_next_partition": (i32, [vp, vp]),
            "spice_balance_ranges": (i32, [vp, i64, i32, vp]),
            "spice_population_range": (i32, [vp, i32, C.POINTER(i64), C.POINTER(i64)]),
            "spice_connect_fixed_probability": (i32, [vp, vp, i32, i32, C.c_double, C.c_float, vp, C.POINTER(i32)]),
            "spice_connect_fixed_probability_fast": (i32, [vp, vp, i32, i32, C.c_double, C.c_float, vp, C.POINTER(i32)]),
            "spice_connect_adj_list": (i32, [vp, vp, i32, i32, vp, vp, i64, C.c_float, vp, C.POINTER(i32)]),
            "spice_connection_csr": (i32, [vp, i32, C.POINTER(i64), vp, vp]),
            "spice_connection_synapses": (i32, [vp, i32, vp, i64]),
            "spice_run": (i32, [vp, i64]),
            "spice_sync": (i32, [vp]),
            "spice_time": (i64, [vp]),
            "spice_spikes": (i32, [vp, i32, i64, C.POINTER(vp), C.POINTER(i64)]),
            "spice_neurons": (i32, [vp, i32, vp, i64]),
            "spice_set_neurons": (i32, [vp, i32, vp, i64]),
            "spice_raster_enable": (i32, [vp, i32]),
            "spice_raster_size": (i32, [vp, i64, C.POINTER(i64), C.POINTER(i64)]),
            "spice_raster_read": (i32, [vp, i64, vp, vp]),
            "spice_stats": (i32, [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]),
            "spice_profile_enable": (i32, [vp, i32]),
            "spice_windows_run": (i64, [vp]),
            "spice_profile_read": (i32, [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                         C.POINTER(i64)]),
            "spice_ctx_finalize": (i32, [vp]),
            "spice_ctx_peer_handle": (i32, [vp, vp, C.POINTER(i64)]),
            "spice_ctx_set_peers": (i32, [vp, vp, i64]),
            "spice_fixed_probability_max_degree": (i64, [i64, C.c_double]),
            "spice_fixed_probability_generate": (i32, [i32, i64, i64, C.c_double, C.c_uint64, C.c_uint64, i64, i64,
                                                       C.POINTER(vp)]),
            "spice_adjacency_edges": (i64, [vp]),
            "spice_adjacency_offsets_dev": (vp, [vp]),
            "spice_adjacency_neighbors_dev": (vp, [vp]),
            "spice_adjacency_copy": (i32, [vp, vp, vp]),
            "spice_adjacency_copy_range": (i32, [vp, i64, i64, vp]),
            "spice_adj_list_generate": (i32, [i32, vp, vp, i64, i64, i64, i64, i64, C.POINTER(vp)]),
            "spice_fixed_probability_generate_fast": (i32, [i32, i64, i64, C.c_double, C.c_uint64, C.c_uint64, i64, i64, C.POINTER(vp)]),
            "spice_adjacency_timing": (i32, [vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(i64)]),
            "spice_adjacency_destroy": (i32, [vp]),
            "spice_seed_seq": (None, [vp, i32, vp]),
            "spice_seed_next": (None, [vp]),
            "spice_builtin_neuron": (vp, [C.c_char_p]),
            "spice_builtin_synapse": (vp, [C.c_char_p]),
            "spice_selftest_libm": (i32, [i32, i32, vp, vp, vp, i64]),
            "spice_device_check": (i32, [i32]),
            "spice_fnv1a64": (C.c_uint64, [vp, i64]),
            "spice_version": (C.c_char_p, []),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# ---- built-in sample models (the same structs as spice/models/*.h) -----------------------------
# name -> (functor packer, neuron dtype)
NEURON_MODELS = {
    "brunel.poisson": (lambda **kw: b"\0", None),
    "brunel.lif": (lambda **kw: b"\0", np.dtype([("V", np.float32), ("Twait", np.int32)])),
    "vogels.lif": (lambda **kw: b"\0", np.dtype([("V", np.float32), ("Gex", np.float32), ("Gin", np.float32),
                                                 ("Twait", np.int32)])),
}
SYNAPSE_MODELS = {
    "brunel.fixed_weight": lambda weight: struct.pack("<f", np.float32(weight)),
    "vogels.excitatory": lambda weight: struct.pack("<f", np.float32(weight)),
    "vogels.inhibitory": lambda weight: struct.pack("<f", np.float32(weight)),
    "brunel+.plastic": lambda: b"\0",
}
# per-synapse state of the stateful synapse models (csr<T>::_edges, csr.h:99)
SYNAPSE_STATE = {
    "brunel+.plastic": np.dtype([("W", np.float32), ("Zpre", np.float32), ("Zpost", np.float32)]),
}


class fixed_probability:
    """spice::fixed_probability (spice/include/spice/topology.h:48-58)."""

    def __init__(self, p: float, fast: bool = False):
        self.p = float(p)
        self.fast = bool(fast)  # this backend's counter-based generator: not the reference's matrix for the same seed


class adj_list:
    """spice::adj_list (spice/include/spice/topology.h:37-46)."""

    def __init__(self, src=None, dst=None):
        """Empty (filled with connect(), as the reference's), or from two index arrays."""
        self.src = [] if src is None else list(np.asarray(src).tolist())
        self.dst = [] if dst is None else list(np.asarray(dst).tolist())

    def connect(self, src: int, dst: int):
        self.src.append(src)
        self.dst.append(dst)


def seed_seq(words, increments=0):
    """util::seed_seq{words...} advanced `increments` times -> (lo, hi)  (random.h:143-175)."""
    w = np.asarray(words, np.uint32)
    out = np.zeros(2, np.uint64)
    lib().spice_seed_seq(_ptr(w), len(w), _ptr(out))
    for _ in range(increments):
        lib().spice_seed_next(_ptr(out))
    return int(out[0]), int(out[1])


HOST_UPDATE_FN = C.CFUNCTYPE(C.c_int64, C.c_void_p, C.c_float, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_int64,
                             C.POINTER(C.c_int64))


class Population:
    """Handle returned by snn.add_population (reference: detail::neuron_population<Neur>*)."""

    def __init__(self, net: "snn", index: int, model: str, size: int):
        self.net, self.index, self.model, self._size = net, index, model, size
        self.dtype = NEURON_MODELS[model][1] if model in NEURON_MODELS else None

    def size(self) -> int:
        return self._size

    def range(self):
        lo, hi = C.c_int64(), C.c_int64()
        self.net._check(lib().spice_population_range(self.net._h, self.index, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def spikes(self, age: int = 0) -> np.ndarray:
        """NeuronPopulation::spikes(age) (neuron_population.h:147-153)."""
        p, n = C.c_void_p(), C.c_int64()
        self.net._check(lib().spice_spikes(self.net._h, self.index, age, C.byref(p), C.byref(n)))
        if n.value == 0:
            return np.zeros(0, np.int32)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int32)), shape=(n.value,)).copy()

    def get_neurons(self) -> np.ndarray:
        """neuron_population::get_neurons() (neuron_population.h:142-145); this rank's slice."""
        if self.dtype is None:
            raise SpiceError(1, "Can only return collections of stateful neurons.")
        lo, hi = self.range()
        out = np.zeros(hi - lo, self.dtype)
        self.net._check(lib().spice_neurons(self.net._h, self.index, _ptr(out), out.nbytes))
        return out

    def set_neurons(self, values: np.ndarray):
        values = np.ascontiguousarray(values, self.dtype)
        self.net._check(lib().spice_set_neurons(self.net._h, self.index, _ptr(values), values.nbytes))


class snn:
    """spice::snn (spice/include/spice/snn.h:16-75) on one B200 (or one rank of several)."""

    def __init__(self, dt, max_delay, seed=(1337,), device=0, rank=0, world=1, mode=MODE_DETERMINISTIC):
        L = lib()
        words = np.asarray(seed, np.uint32)
        h = C.c_void_p()
        rc = L.spice_ctx_create(C.byref(h), device, np.float32(dt), np.float32(max_delay), _ptr(words), len(words), rank,
                                world, mode)
        if rc != 0:
            raise SpiceError(rc, L.spice_last_error(None).decode())
        self._h = h
        self.populations: list[Population] = []
        self.connections = []
        self._keepalive = []  # ctypes callbacks of host-fed populations
        self.rank, self.world = rank, world

    def close(self):
        if getattr(self, "_h", None):
            lib().spice_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise SpiceError(rc, lib().spice_last_error(self._h).decode())

    def set_stream(self, cuda_stream: int):
        self._check(lib().spice_ctx_set_stream(self._h, C.c_void_p(cuda_stream)))

    def add_population(self, model: str, size: int, bounds=None, **params) -> Population:
        """bounds (optional, world + 1 entries): the ranks' target ranges instead of equal widths (balance_ranges())."""
        ops = lib().spice_builtin_neuron(model.encode())
        if not ops:
            raise SpiceError(3, f"unknown neuron model {model!r}")
        functor = NEURON_MODELS[model][0](**params)
        idx = C.c_int()
        if bounds is not None:
            b = np.ascontiguousarray(bounds, np.int64)
            self._check(lib().spice_set_next_partition(self._h, _ptr(b)))
        self._check(lib().spice_add_population(self._h, ops, size, functor, C.byref(idx)))
        pop = Population(self, idx.value, model, size)
        self.populations.append(pop)
        return pop

    def add_host_population(self, size: int, update) -> Population:
        """A population whose spikes come from the host (a neuron with a per-population update(),
        concepts.h:46-57): `update(dt)` is called once per step, in step order, when the step is enqueued,
        and returns the ids of the neurons that fire."""
        def trampoline(_user, dt, _lo, _hi, _off, ids_out, capacity, draws_out):
            try:
                ids = np.asarray(update(dt), np.int32).ravel()
                if len(ids) > capacity:
                    return -1
                C.memmove(ids_out, ids.ctypes.data, ids.nbytes)
                draws_out[0] = 0
                return len(ids)
            except Exception:  # noqa: BLE001 - reported as a failed step by the runtime
                return -1

        cb = HOST_UPDATE_FN(trampoline)
        self._keepalive.append(cb)
        idx = C.c_int()
        self._check(lib().spice_add_host_population(self._h, size, C.cast(cb, C.c_void_p), None, C.byref(idx)))
        pop = Population(self, idx.value, "host", size)
        self.populations.append(pop)
        return pop

    def connect(self, synapse: str, source: Population, target: Population, topology, delay, **params) -> int:
        ops = lib().spice_builtin_synapse(synapse.encode())
        if not ops:
            raise SpiceError(3, f"unknown synapse model {synapse!r}")
        functor = SYNAPSE_MODELS[synapse](**params)
        idx = C.c_int()
        if isinstance(topology, fixed_probability):
            fn = lib().spice_connect_fixed_probability_fast if topology.fast else lib().spice_connect_fixed_probability
            self._check(fn(self._h, ops, source.index, target.index, topology.p, np.float32(delay), functor, C.byref(idx)))
        elif isinstance(topology, adj_list):
            s = np.asarray(topology.src, np.int32)
            d = np.asarray(topology.dst, np.int32)
            self._check(lib().spice_connect_adj_list(self._h, ops, source.index, target.index, _ptr(s), _ptr(d), len(s),
                                                     np.float32(delay), functor, C.byref(idx)))
        else:
            raise TypeError("topology must be fixed_probability or adj_list")
        self.connections.append((synapse, source.index, target.index))
        return idx.value

    def connection_csr(self, conn: int):
        n = C.c_int64()
        self._check(lib().spice_connection_csr(self._h, conn, C.byref(n), None, None))
        src_size = self.populations[self.connections[conn][1]].size()
        off = np.zeros(src_size + 1, np.int64)
        nb = np.zeros(max(n.value, 1), np.int32)
        self._check(lib().spice_connection_csr(self._h, conn, C.byref(n), _ptr(off), _ptr(nb)))
        return off, nb[: n.value]

    def connection_synapses(self, conn: int) -> np.ndarray:
        """csr<T>::_edges of a stateful connection (csr.h:99), parallel to connection_csr()[1]."""
        dt = SYNAPSE_STATE[self.connections[conn][0]]
        out = np.zeros(self.connection_edges(conn), dt)
        self._check(lib().spice_connection_synapses(self._h, conn, _ptr(out), out.nbytes))
        return out

    def connection_edges(self, conn: int) -> int:
        n = C.c_int64()
        self._check(lib().spice_connection_csr(self._h, conn, C.byref(n), None, None))
        return n.value

    def step(self, n: int = 1):
        """snn::step() n times (spice/src/snn.cpp:7-28).  Asynchronous; readouts synchronise."""
        self._check(lib().spice_run(self._h, n))

    def sync(self):
        self._check(lib().spice_sync(self._h))

    def time(self) -> int:
        return int(lib().spice_time(self._h))

    def spikes(self, i: int, age: int = 0) -> np.ndarray:
        """Convenience named by the north star: spikes of the i-th population added."""
        return self.populations[i].spikes(age)

    def raster_enable(self, on=True):
        self._check(lib().spice_raster_enable(self._h, int(on)))

    def raster_read(self, steps: int = 0):
        """-> (counts[steps, npops], ascending ids concatenated in (step, pop) order) of the first `steps`
        unread steps (0: every step issued so far); waits only for those steps and frees their log space."""
        n, nids = C.c_int64(), C.c_int64()
        self._check(lib().spice_raster_size(self._h, steps, C.byref(n), C.byref(nids)))
        counts = np.empty((n.value, len(self.populations)), np.int64)
        ids = np.empty(max(nids.value, 1), np.int32)
        self._check(lib().spice_raster_read(self._h, n.value, _ptr(counts), _ptr(ids)))
        return counts, ids[: nids.value]

    def seed(self):
        """snn::_seed (snn.h:69): (lo, hi) of the seed the next `seed++` hands out."""
        out = (C.c_uint64 * 2)()
        self._check(lib().spice_ctx_seed(self._h, out))
        return int(out[0]), int(out[1])

    def stats(self):
        ev, sp, kl = C.c_int64(), C.c_int64(), C.c_int64()
        self._check(lib().spice_stats(self._h, C.byref(ev), C.byref(sp), C.byref(kl)))
        return dict(synaptic_events=ev.value, spikes_delivered=sp.value, kernel_launches=kl.value)

    def windows_run(self) -> int:
        return int(lib().spice_windows_run(self._h))

    def profile_enable(self, on=True, every=1):
        """Phase events around every `every`-th window (1: all of them)."""
        self._check(lib().spice_profile_enable(self._h, (max(int(every), 1) if on else 0)))

    def profile_read(self):
        u, d, x, w = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
        self._check(lib().spice_profile_read(self._h, C.byref(u), C.byref(d), C.byref(x), C.byref(w)))
        return dict(update_ms=u.value, deliver_ms=d.value, exchange_ms=x.value, windows=w.value)

    # multi-GPU plumbing (one process per GPU; handles are all-gathered by the caller)
    def finalize(self):
        self._check(lib().spice_ctx_finalize(self._h))

    def peer_handle(self) -> bytes:
        n = C.c_int64()
        self._check(lib().spice_ctx_peer_handle(self._h, None, C.byref(n)))
        buf = C.create_string_buffer(n.value)
        self._check(lib().spice_ctx_peer_handle(self._h, buf, C.byref(n)))
        return buf.raw

    def set_peers(self, handles: list[bytes]):
        blob = b"".join(handles)
        self._check(lib().spice_ctx_set_peers(self._h, blob, len(handles[0])))


def fnv1a64(a) -> str:
    """FNV-1a (64 bit) of an array's bytes as 16 hex digits — the digest the golden fixtures use."""
    a = np.ascontiguousarray(a)
    return f"{int(lib().spice_fnv1a64(_ptr(a), a.nbytes)):016x}"


class Adjacency:
    """A generated adjacency that stays on the device (bench/connectivity sizes whose arrays do not fit the host):
    offsets() and rows(lo, hi) copy what is asked for."""

    def __init__(self, src, dst, p, seed=(1337,), increments=0, device=0, col_lo=0, col_hi=None, fast=False):
        L = lib()
        lo, hi = seed_seq(seed, increments)
        self.src, self.dst = src, dst
        self.col_lo, self.col_hi = col_lo, dst if col_hi is None else col_hi
        self.h = C.c_void_p()
        gen = L.spice_fixed_probability_generate_fast if fast else L.spice_fixed_probability_generate
        rc = gen(device, src, dst, p, lo, hi, self.col_lo, self.col_hi, C.byref(self.h))
        if rc != 0:
            raise SpiceError(rc, L.spice_last_error(None).decode())
        self.edges = int(L.spice_adjacency_edges(self.h))
        total, rows = C.c_float(), C.c_float()
        draws = C.c_int64()
        L.spice_adjacency_timing(self.h, C.byref(total), C.byref(rows), C.byref(draws))
        self.total_ms, self.rows_ms, self.draws = total.value, rows.value, draws.value
        self._off = None

    def offsets(self):
        if self._off is None:
            off = np.zeros(self.src + 1, np.int64)
            if lib().spice_adjacency_copy(self.h, _ptr(off), None) != 0:
                raise SpiceError(2, "adjacency copy failed")
            self._off = off
        return self._off

    def rows(self, lo, hi):
        """neighbors of rows [lo, hi) (local columns), one flat array."""
        off = self.offsets()
        out = np.zeros(max(int(off[hi] - off[lo]), 1), np.int32)
        if lib().spice_adjacency_copy_range(self.h, int(off[lo]), int(off[hi]), _ptr(out)) != 0:
            raise SpiceError(2, "adjacency copy failed")
        return out[: int(off[hi] - off[lo])]

    def close(self):
        if self.h:
            lib().spice_adjacency_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def generate_adj_list(edges_src, edges_dst, src_count, dst_count, device=0, col_lo=0, col_hi=None, copy=True):
    """adj_list::generate on the GPU (spice/src/topology.cpp:56-71): (src, dst) pairs -> CSR sorted by (src, dst) -> dict."""
    L = lib()
    es = np.ascontiguousarray(edges_src, np.int32)
    ed = np.ascontiguousarray(edges_dst, np.int32)
    assert es.shape == ed.shape and es.ndim == 1
    col_hi = dst_count if col_hi is None else col_hi
    h = C.c_void_p()
    rc = L.spice_adj_list_generate(device, _ptr(es), _ptr(ed), es.size, src_count, dst_count, col_lo, col_hi, C.byref(h))
    if rc != 0:
        raise SpiceError(rc, L.spice_last_error(None).decode())
    try:
        e = int(L.spice_adjacency_edges(h))
        total, rows = C.c_float(), C.c_float()
        draws = C.c_int64()
        L.spice_adjacency_timing(h, C.byref(total), C.byref(rows), C.byref(draws))
        out = dict(edges=e, total_ms=total.value)
        if copy:
            off = np.zeros(src_count + 1, np.int64)
            nb = np.zeros(max(e, 1), np.int32)
            if L.spice_adjacency_copy(h, _ptr(off), _ptr(nb)) != 0:
                raise SpiceError(2, "adjacency copy failed")
            out.update(offsets=off, neighbors=nb[:e])
        return out
    finally:
        L.spice_adjacency_destroy(h)


def balance_ranges(in_degree, world: int) -> np.ndarray:
    """Static synapse-count load balancing: target ranges holding about equal sums of in_degree + 1."""
    w = np.ascontiguousarray(in_degree, np.int64)
    out = np.zeros(world + 1, np.int64)
    rc = lib().spice_balance_ranges(_ptr(w), len(w), world, _ptr(out))
    if rc != 0:
        raise SpiceError(rc, "spice_balance_ranges: invalid argument")
    return out


def generate_fixed_probability(src, dst, p, seed=(1337,), increments=0, device=0, col_lo=0, col_hi=None, copy=True, fast=False):
    """fixed_probability::generate on the GPU (spice/src/topology.cpp:80-112) -> dict.  fast=True: the counter-based
    generator (spice_fixed_probability_generate_fast), not the reference's matrix."""
    L = lib()
    lo, hi = seed_seq(seed, increments)
    col_hi = dst if col_hi is None else col_hi
    h = C.c_void_p()
    gen = L.spice_fixed_probability_generate_fast if fast else L.spice_fixed_probability_generate
    rc = gen(device, src, dst, p, lo, hi, col_lo, col_hi, C.byref(h))
    if rc != 0:
        raise SpiceError(rc, L.spice_last_error(None).decode())
    try:
        e = int(L.spice_adjacency_edges(h))
        total, rows = C.c_float(), C.c_float()
        draws = C.c_int64()
        L.spice_adjacency_timing(h, C.byref(total), C.byref(rows), C.byref(draws))
        out = dict(edges=e, total_ms=total.value, rows_ms=rows.value, draws=draws.value)
        if copy:
            off = np.zeros(src + 1, np.int64)
            nb = np.zeros(max(e, 1), np.int32)
            rc = L.spice_adjacency_copy(h, _ptr(off), _ptr(nb))
            if rc != 0:
                raise SpiceError(rc, "adjacency copy failed")
            out.update(offsets=off, neighbors=nb[:e])
        return out
    finally:
        L.spice_adjacency_destroy(h)
