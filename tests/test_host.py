"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol the public
header declares, the host RNG / jump-ahead / glibc-log restatements agree with the oracle and the
golden vectors.  No compute call is made (no GPU here)."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
hx = lambda v: f"{int(v):016x}"


@pytest.fixture(scope="module")
def sp():
    import spice2_b200 as sp

    sp.lib()
    return sp


def test_library_exports_every_declared_symbol(sp):
    header = (ROOT / "include" / "spice_b200.h").read_text()
    declared = set(re.findall(r"SPICE_API [^;(]*?\b(spice_\w+)\(", header))
    assert len(declared) >= 35
    so = ROOT / "spice2_b200" / "libspice_b200.so"
    out = subprocess.run(["nm", "-D", "--defined-only", str(so)], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (spice_\w+)", out))
    assert declared <= exported, declared - exported
    L = C.CDLL(str(so))
    for name in declared:
        getattr(L, name)


def test_library_is_sm100a_cuda(sp):
    so = ROOT / "spice2_b200" / "libspice_b200.so"
    out = subprocess.run(["cuobjdump", "-lelf", str(so)], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert b"spice2_b200" in sp.lib().spice_version()


def test_no_device_fails_loudly(sp):
    """No CPU fallback: without a CUDA device the computing entry points refuse."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(sp.SpiceError, match="no CPU fallback|no CUDA device"):
        sp.snn(1e-4, 15e-4)
    with pytest.raises(sp.SpiceError):
        sp.generate_fixed_probability(10, 10, 0.5)


def test_product_never_imports_oracle():
    for p in (ROOT / "spice2_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h", ".cpp") and p.is_file():
            text = p.read_text()
            for needle in ("liboracle", "oracle_lib", "spice_oracle", "oracle/"):
                assert needle not in text, f"{p} mentions {needle}"


def test_host_seed_seq_matches_golden(sp, golden):
    for k, want in golden["seed_1337"].items():
        assert [hx(x) for x in sp.seed_seq([1337], int(k))] == want
    for il, want in golden["seed_il"].items():
        assert [hx(x) for x in sp.seed_seq([int(x) for x in il.split(",")])] == want


def test_max_degree_matches_golden(sp, golden):
    for g in golden["fixed_probability"]:
        if g["src"] > 0:
            assert sp.lib().spice_fixed_probability_max_degree(g["dst"], g["p"]) * g["src"] == g["capacity"]


@pytest.fixture(scope="module")
def hosttool(tmp_path_factory):
    """Small host program exercising the header-only restatements (glibc log, jump-ahead)."""
    d = tmp_path_factory.mktemp("hosttool")
    src = d / "t.cpp"
    src.write_text(r'''
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include "spice/detail/glibc_log.h"
#include "spice/detail/glibc_expf.h"
#include "spice/util/random.h"
#include "spice/util/numeric.h"
using namespace spice::util;
int main(int argc, char** argv) {
	if (!strcmp(argv[1], "log")) {   // stdin: doubles as hex bit patterns; stdout: log bits
		unsigned long long u;
		while (scanf("%llx", &u) == 1) printf("%016llx\n", (unsigned long long)spice::detail::glibc::double_to_bits(spice::detail::glibc::log(spice::detail::glibc::bits_to_double(u))));
	} else if (!strcmp(argv[1], "expf")) { // stdin: floats as hex bit patterns; stdout: expf bits
		unsigned u;
		while (scanf("%x", &u) == 1) { float x; memcpy(&x, &u, 4); float y = spice::detail::glibc::expf_restated(x); memcpy(&u, &y, 4); printf("%08x\n", u); }
	} else if (!strcmp(argv[1], "jump")) { // args: lo hi k -> state after k steps via jump polynomial
		xoroshiro64_128p s(strtoull(argv[2], 0, 16), strtoull(argv[3], 0, 16));
		auto r = jump::apply(jump::xpow(strtoull(argv[4], 0, 10)), s);
		printf("%016llx %016llx\n", (unsigned long long)r.s0, (unsigned long long)r.s1);
	} else if (!strcmp(argv[1], "dist")) { // args: kind a b count ; seed {7, 9} advanced 2 times -> values as double bit patterns
		int kind = atoi(argv[2]); double a = atof(argv[3]), b = atof(argv[4]); int n = atoi(argv[5]);
		seed_seq seq{7, 9}; seq++; seq++;
		xoroshiro32_128p r32(seq); xoroshiro64_128p r64(seq);
		auto put = [](double v) { unsigned long long u; memcpy(&u, &v, 8); printf("%016llx\n", u); };
		if (kind == 0) for (int i = 0; i < n; i++) put((double)r32());
		if (kind == 1) for (int i = 0; i < n; i++) put(generate_canonical<double>(r32));
		if (kind == 2) { normal_distribution<double> d(a, b); for (int i = 0; i < n; i++) put(d(r64)); }
		if (kind == 3) { normal_distribution<float> d((float)a, (float)b); for (int i = 0; i < n; i++) put(d(r64)); }
		if (kind == 4) { binomial_distribution<Int> d((Int)a, b); for (int i = 0; i < n; i++) put((double)d(r64)); }
		if (kind == 5) { exponential_distribution<float> d((float)a); for (int i = 0; i < n; i++) put(d(r64)); }
		if (kind == 6) { unsigned w[2] = {7, 9}; seed_seq s2{std::seed_seq(w, w + 2)}; unsigned o[8]; s2.generate(o, o + 8); for (int i = 0; i < n && i < 8; i++) put((double)o[i]); }
	} else if (!strcmp(argv[1], "kahan")) {
		kahan_sum<float> k; int n = atoi(argv[2]);
		for (int i = 0; i < n; i++) { float d = k += 1e-4f; if (k >= 1) k.reset(); unsigned b; memcpy(&b, &d, 4); printf("%08x\n", b); }
	}
	return 0;
}
''')
    exe = d / "t"
    inc = ROOT / "spice2_b200" / "csrc" / "include"
    subprocess.run(["g++", "-std=c++20", "-O2", "-ffp-contract=off", "-mfma", f"-I{inc}", str(src),
                    str(ROOT / "spice2_b200" / "csrc" / "jump.cpp"), "-o", str(exe)], check=True)
    return exe


def test_glibc_log_restatement_matches_golden_pins(hosttool):
    z = np.load(ROOT / "tests" / "golden" / "libm_pins.npz")
    x, y = z["log_x"], z["log_y"]
    inp = "\n".join(f"{int(v):x}" for v in x.view(np.uint64))
    out = subprocess.run([str(hosttool), "log"], input=inp, capture_output=True, text=True, check=True).stdout.split()
    got = np.array([int(v, 16) for v in out], np.uint64)
    assert np.array_equal(got, y.view(np.uint64))


def test_glibc_expf_restatement_matches_golden_pins(hosttool):
    z = np.load(ROOT / "tests" / "golden" / "libm_pins.npz")
    x, y = z["expf_x"].astype(np.float32), z["expf_y"].astype(np.float32)
    inp = "\n".join(f"{int(v):x}" for v in x.view(np.uint32))
    out = subprocess.run([str(hosttool), "expf"], input=inp, capture_output=True, text=True, check=True).stdout.split()
    got = np.array([int(v, 16) for v in out], np.uint32)
    assert np.array_equal(got, y.view(np.uint32))


def test_jump_ahead_matches_sequential(hosttool, orc):
    seed = orc.seed_seq([1337], 3)
    for k in (0, 1, 2, 127, 128, 129, 1000, 65537, 1234567):
        out = subprocess.run([str(hosttool), "jump", hx(seed.lo), hx(seed.hi), str(k)], capture_output=True, text=True,
                             check=True).stdout.split()
        assert (int(out[0], 16), int(out[1], 16)) == orc.state_at(seed, k), k


def test_host_kahan_dt_matches_oracle(hosttool, orc):
    out = subprocess.run([str(hosttool), "kahan", "12000"], capture_output=True, text=True, check=True).stdout.split()
    got = np.array([int(v, 16) for v in out], np.uint32)
    assert np.array_equal(got, orc.kahan_dt(np.float32(1e-4), 12000).view(np.uint32))


def test_topology_header_edge_stream(tmp_path):
    """spice/topology.h on the host: the reference's CSR cases (test/detail/csr.cpp:17-61: Empty, Custom) through
    Topology::generate(offsets, neighbors, seed), a user-defined Topology writing an edge_stream (topology.h:11-22,
    topology.cpp:12-55), and the preconditions of edge_stream / the Topology base class."""
    src = tmp_path / "topo.cpp"
    src.write_text(r'''
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <vector>
#include "spice/topology.h"
using namespace spice;
#define CHECK(c) do { if (!(c)) { std::printf("line %d: %s\n", __LINE__, #c); return 1; } } while (0)
struct ring : Topology {       // i -> i + 1 (mod n), and 0 -> every third neuron
	Int size() const override { return src_count + dst_count; }
	using Topology::generate;
	void generate(edge_stream& s, util::seed_seq const& seed) override {
		util::xoroshiro64_128p rng(seed);
		draw = rng();
		for (Int i = 0; i < src_count; i++) {
			if (i == 0)
				for (Int d = 0; d < dst_count; d += 3)
					s << std::pair{Int32(0), Int32(d)};
			if (i != 2)            // row 2 stays empty
				s << std::pair{Int32(i), Int32((i + 1) % dst_count)};
		}
	}
	UInt draw = 0;
};
struct nothing : Topology { Int size() const override { return 0; } };
template <class F> bool throws(F f) { try { f(); } catch (std::logic_error const&) { return true; } return false; }
int main() {
	{ // CSR.Empty
		adj_list adj; adj(1, 0);
		std::vector<Int> off(2); std::vector<Int32> nb; // zero-filled, as csr's constructor sizes them (csr.h:70-71)
		adj.generate(off, nb, util::seed_seq{1337});
		CHECK(off[0] == 0 && off[1] == 0);
	}
	{ // CSR.Custom: unsorted input, empty middle row
		adj_list adj;
		CHECK(adj.size() == 0);
		adj.connect(2, 2); adj.connect(0, 9); adj.connect(0, 0); adj.connect(2, 1); adj.connect(0, 7); adj.connect(0, 2);
		adj(3, 10);
		std::vector<Int> off(4, -1); std::vector<Int32> nb(adj.size(), -1);
		adj.generate(off, nb, util::seed_seq{1337});
		CHECK((off == std::vector<Int>{0, 4, 4, 6}));
		CHECK((nb == std::vector<Int32>{0, 2, 7, 9, 1, 2}));
	}
	{ // user-defined topology; trailing empty rows are closed by flush()
		ring r; r(6, 5);
		std::vector<Int> off(7, -1); std::vector<Int32> nb(r.size(), -1);
		r.generate(off, nb, util::seed_seq{1337});
		CHECK((off == std::vector<Int>{0, 3, 4, 4, 5, 6, 7}));
		CHECK((std::vector<Int32>(nb.begin(), nb.begin() + 7) == std::vector<Int32>{0, 3, 1, 2, 4, 0, 1}));
		CHECK(r.draw == 0xa349d5b00b80b8e0ull); // first output of xoroshiro64_128p(seed_seq{1337}) (SURVEY 8c)
	}
	{ // preconditions
		nothing n; n(1, 1);
		std::vector<Int> off(2); std::vector<Int32> nb(1);
		CHECK(throws([&] { n.generate(off, nb, util::seed_seq{1}); }));            // generate(edge_stream) not implemented
		ring r; r(6, 5);
		std::vector<Int> small(6);  std::vector<Int32> big(64);
		CHECK(throws([&] { r.generate(small, big, util::seed_seq{1}); }));        // offsets.size() > src_count
		std::vector<Int> ok(7); std::vector<Int32> tiny(3);
		CHECK(throws([&] { r.generate(ok, tiny, util::seed_seq{1}); }));          // neighbors.size() >= size()
		edge_stream es(ok, big);
		CHECK(throws([&] { es << std::pair{Int32(7), Int32(0)}; }));               // src beyond the offsets
		CHECK(throws([&] { Topology& t = r; t(-1, 3); }));
	}
	std::puts("ok");
	return 0;
}
''')
    exe = tmp_path / "topo"
    inc = ROOT / "spice2_b200" / "csrc" / "include"
    subprocess.run(["g++", "-std=c++20", "-O1", "-Wall", f"-I{inc}", str(src), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stdout + r.stderr


def test_delivery_tickets_cover_every_unit(tmp_path):
    """csrc/deliver_plan.h (the ticket arithmetic of the delivery kernel, shared with it): the static tickets of a
    grid's CTAs plus the dynamically claimed ones visit every unit exactly once whatever the grid size, and
    locate_unit() inverts unit = tile_prefix[c] * nsteps + s * tiles[c] + k."""
    src = tmp_path / "plan.cpp"
    src.write_text(r"""
#include <cstdio>
#include <random>
#include <vector>
#include "deliver_plan.h"
using namespace spice::deliver;
#define CHECK(c) do { if (!(c)) { std::printf("line %d: %s\n", __LINE__, #c); return 1; } } while (0)
int main() {
	std::mt19937 g(7);
	for (int trial = 0; trial < 300; trial++) {
		int const nconns = 1 + g() % 7, nsteps = 1 + g() % 15;
		std::vector<int> prefix(nconns + 1, 0);
		for (int c = 0; c < nconns; c++)
			prefix[c + 1] = prefix[c] + 1 + g() % 60;
		unsigned const units = static_cast<unsigned>(prefix[nconns]) * nsteps;
		// locate_unit inverts the numbering
		unsigned u = 0;
		for (int c = 0; c < nconns; c++)
			for (int s = 0; s < nsteps; s++)
				for (int k = 0; k < prefix[c + 1] - prefix[c]; k++, u++) {
					unit_pos const p = locate_unit(u, prefix, nconns, nsteps);
					CHECK(p.c == c && p.s == s && p.k == k);
				}
		CHECK(u == units);
		// static + dynamic tickets: a partition of [0, units) for any grid
		unsigned const grid = 1 + g() % 400;
		std::vector<int> seen(units, 0);
		for (unsigned cta = 0; cta < grid; cta++)
			for (unsigned j = 0; j < static_cast<unsigned>(kStaticUnits); j++) {
				unsigned const t = static_ticket(cta, grid, j);
				if (t < units)
					seen[t]++;
			}
		for (unsigned claimed = 0; dynamic_ticket(grid, claimed) < units; claimed++)
			seen[dynamic_ticket(grid, claimed)]++;
		for (unsigned i = 0; i < units; i++)
			CHECK(seen[i] == 1);
		CHECK(kTicketRing > kStaticUnits);
	}
	std::puts("ok");
	return 0;
}
""")
    exe = tmp_path / "plan"
    inc = ROOT / "spice2_b200" / "csrc"
    subprocess.run(["g++", "-std=c++20", "-O1", "-Wall", f"-I{inc}", str(src), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stdout + r.stderr


def test_counter_address_rotation(tmp_path):
    """csrc/deliver_plan.h rotw_fwd / rotw_inv (the word of a target's second u32 counter in the delivery stream):
    a bijection on every 32-word row of a tile, and of the targets that share a bank with a given target in array A
    (one per row), few share its bank in array B as well — what makes pack_runs' 2-choice balancing effective."""
    src = tmp_path / "rot.cpp"
    src.write_text(r"""
#include <cstdio>
#include <set>
#include "deliver_plan.h"
using namespace spice::deliver;
#define CHECK(c) do { if (!(c)) { std::printf("line %d: %s (t %d)\n", __LINE__, #c, t); return 1; } } while (0)
int main() {
	int const cap = 5120;
	std::set<int> seen;
	for (int t = 0; t < cap; t++) {
		int const u = rotw_fwd(t);
		CHECK(u >= 0 && u < cap);
		CHECK((u >> 5) == (t >> 5)); // stays in its 128-byte row
		CHECK(rotw_inv(u) == t);
		CHECK(seen.insert(u).second);
		CHECK((u & 31) == ((t + rotw_amount(t >> 5)) & 31));
	}
	// rows r and r + 1 .. r + 31 never rotate by the same amount; over the 160 rows of the largest tile at most 5 do
	for (int r = 0; r < 160; r++) {
		int same = 0;
		for (int q = 0; q < 160; q++)
			same += rotw_amount(q) == rotw_amount(r);
		if (same > 5) { std::printf("row %d shares its rotation with %d rows\n", r, same); return 1; }
		for (int q = r + 1; q < r + 32 && q < 160; q++)
			if ((q >> 5) == (r >> 5) && rotw_amount(q) == rotw_amount(r)) { std::printf("rows %d and %d\n", r, q); return 1; }
	}
	std::puts("ok");
	return 0;
}
""")
    exe = tmp_path / "rot"
    inc = ROOT / "spice2_b200" / "csrc"
    subprocess.run(["g++", "-std=c++20", "-O1", "-Wall", f"-I{inc}", str(src), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stdout + r.stderr


def test_balance_ranges_host_arithmetic(sp):
    """spice_balance_ranges (static synapse-count load balancing) is host arithmetic: equal in-degrees give
    equal widths, a skewed histogram is cut where its prefix sum crosses r / world of the total."""
    assert sp.balance_ranges(np.full(1000, 7), 4).tolist() == [0, 250, 500, 750, 1000]
    assert sp.balance_ranges(np.zeros(0, np.int64), 3).tolist() == [0, 0, 0, 0]
    w = np.zeros(100, np.int64)
    w[:10] = 1000  # ten heavy targets: each is ~1/10 of the work
    b = sp.balance_ranges(w, 2)
    assert b.tolist() == [0, 5, 100]
    rng = np.random.default_rng(1)
    w = (rng.pareto(1.5, 50000) * 100).astype(np.int64)
    for world in (2, 3, 8):
        b = sp.balance_ranges(w, world)
        assert b[0] == 0 and b[-1] == len(w) and np.all(np.diff(b) >= 0)
        load = np.add.reduceat(w + 1, b[:-1])
        assert load.max() <= (w + 1).sum() / world + (w.max() + 1)
    with pytest.raises(sp.SpiceError):
        sp.balance_ranges(np.array([1, -2, 3]), 2)


def test_random_header_distributions_match_the_compiled_reference(hosttool):
    """The drop-in random.h's 32-bit engine, two-draw canonical, normal / binomial / float exponential distributions and
    seed_seq(std::seed_seq) against the reference's own header compiled with IEEE flags (oracle/_ref, strict flavour): value
    for value on the same seed (the facade must carry the reference's whole random.h surface)."""
    from oracle_lib import RefShim

    if not RefShim.available("strict"):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    ref = RefShim("strict")
    for kind, a, b, n in [(0, 0, 0, 64), (1, 0, 0, 64), (2, 1.5, 0.25, 101), (3, -2.0, 3.0, 101), (4, 1000, 0.3, 51), (5, 0.7, 0, 64), (6, 0, 0, 8)]:
        out = subprocess.run([str(hosttool), "dist", str(kind), repr(float(a)), repr(float(b)), str(n)], capture_output=True, text=True,
                             check=True).stdout.split()
        got = np.array([int(v, 16) for v in out], np.uint64).view(np.float64)
        want = ref.random_sample(kind, (7, 9), 2, a, b, n)
        assert np.array_equal(got, want), (kind, got[:4], want[:4])
