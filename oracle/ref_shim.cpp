// TEST INFRASTRUCTURE — not product code.
//
// C-ABI shim over the UNMODIFIED reference (denniskb/spice2 @ f5e57eb), compiled from the
// sources where they lie under $(REF) (= /root/reference) by oracle/Makefile into
// oracle/_ref/libspice_ref_{fast,strict}.so.  Nothing from the reference is copied into this
// repository: this file only calls the reference's public API (spice/snn.h, spice/topology.h,
// spice/util/random.h, spice/util/numeric.h) and textually includes the reference's sample
// model definitions (samples/brunel.cpp, samples/brunel+.cpp, samples/vogels.cpp) in place,
// inside namespaces, so that the oracle runs the reference's own functors.
//
// Only tests/, __graft_entry__.smoke() and bench.py's reference / cpu_baseline leg may load
// the resulting library.

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <initializer_list>
#include <iostream>
#include <limits>
#include <memory>
#include <span>
#include <string>
#include <vector>

#include "spice/snn.h"
#include "spice/topology.h"
#include "spice/util/numeric.h"
#include "spice/util/random.h"
#include "spice/util/range.h"

#include "matplot.h" // reference samples/matplot.h (declarations only; the shim defines the sink)

// The samples call pause() and stream into spike_output_stream; the shim never runs their
// main(), but the symbols must resolve.  (The reference's own sink is samples/matplot.cpp,
// which the sample executables built by the Makefile link instead.)
#ifndef REF_SHIM_NO_SINK
void pause(double) {}
struct spike_output_stream::impl {};
spike_output_stream::spike_output_stream(std::string const&, bool) {}
spike_output_stream::~spike_output_stream() = default;
spike_output_stream& spike_output_stream::operator<<(spice::detail::NeuronPopulation const*) {
	return *this;
}
spike_output_stream& spike_output_stream::operator<<(char const) { return *this; }
#endif

// The reference's own model definitions, in place.  All headers they include are already
// included above at global scope (every one is `#pragma once`), so inside the namespace the
// include directives are no-ops and only the sample's own declarations land here.
namespace ref_brunel {
#include "brunel.cpp"
}
namespace ref_brunel_plus {
#include "brunel+.cpp"
}
namespace ref_vogels {
#include "vogels.cpp"
}
namespace ref_sssp {
#include "sssp.cpp"
}

namespace {
using clk = std::chrono::steady_clock;

double seconds(clk::time_point a, clk::time_point b) {
	return std::chrono::duration<double>(b - a).count();
}

spice::util::seed_seq make_seed(std::uint32_t const* il, int n, int increments) {
	// seed_seq only takes an initializer_list; build one per supported arity.
	spice::util::seed_seq s = [&]() -> spice::util::seed_seq {
		switch (n) {
			case 1: return {il[0]};
			case 2: return {il[0], il[1]};
			case 3: return {il[0], il[1], il[2]};
			case 4: return {il[0], il[1], il[2], il[3]};
			default: return {il[0], il[1], il[2], il[3], il[4]};
		}
	}();
	for (int i = 0; i < increments; i++)
		s++;
	return s;
}

struct raster {
	std::int32_t* ids;     // flat spike ids, population-local
	std::int64_t capacity; // capacity of ids
	std::int64_t* counts;  // [steps * npop]
	std::int64_t used = 0;
	bool overflow     = false;

	void push(std::span<std::int32_t const> s, std::int64_t slot) {
		counts[slot] = static_cast<std::int64_t>(s.size());
		if (used + static_cast<std::int64_t>(s.size()) > capacity) {
			overflow = true;
			return;
		}
		std::copy(s.begin(), s.end(), ids + used);
		used += static_cast<std::int64_t>(s.size());
	}
};
}

extern "C" {

// ---- seeds / rng (random.h:143-175, 222-234) -------------------------------------------------
void ref_seed(std::uint32_t const* il, int n, int increments, std::uint64_t out[2]) {
	auto const s = make_seed(il, n, increments).seed();
	out[0]       = s.lo;
	out[1]       = s.hi;
}

void ref_xoroshiro(std::uint32_t const* il, int n, int increments, std::int64_t count,
                   std::uint64_t* out) {
	spice::util::xoroshiro64_128p rng(make_seed(il, n, increments));
	for (std::int64_t i = 0; i < count; i++)
		out[i] = rng();
}

// generate_canonical<float>(rng) and exponential_distribution<double>(scale)(rng) on the same
// stream, so tests can pin the u -> value maps (random.h:236-247, 264-276).
void ref_canonical_float(std::uint32_t const* il, int n, int increments, std::int64_t count,
                         float* out) {
	spice::util::xoroshiro64_128p rng(make_seed(il, n, increments));
	for (std::int64_t i = 0; i < count; i++)
		out[i] = spice::util::generate_canonical<float>(rng);
}

void ref_exponential(std::uint32_t const* il, int n, int increments, double scale,
                     std::int64_t count, double* out) {
	spice::util::xoroshiro64_128p rng(make_seed(il, n, increments));
	spice::util::exponential_distribution<double> d(scale);
	for (std::int64_t i = 0; i < count; i++)
		out[i] = d(rng);
}

// The reference's remaining engines / distributions on one stream (random.h:204-220, 236-330), so the drop-in header's
// versions can be pinned value for value.  kind: 0 xoroshiro32_128p raw draws, 1 generate_canonical<double> from the 32-bit
// engine, 2 normal_distribution<double>(a, b), 3 normal_distribution<float>(a, b), 4 binomial_distribution<Int>(a, b),
// 5 exponential_distribution<float>(a), 6 seed_seq(std::seed_seq{il...})'s seed words.
void ref_random_sample(int kind, std::uint32_t const* il, int n, int increments, double a, double b, std::int64_t count, double* out) {
	auto const seed = make_seed(il, n, increments);
	spice::util::xoroshiro32_128p rng32(seed);
	spice::util::xoroshiro64_128p rng64(seed);
	if (kind == 0)
		for (std::int64_t i = 0; i < count; i++)
			out[i] = static_cast<double>(rng32());
	else if (kind == 1)
		for (std::int64_t i = 0; i < count; i++)
			out[i] = spice::util::generate_canonical<double>(rng32);
	else if (kind == 2) {
		spice::util::normal_distribution<double> d(a, b);
		for (std::int64_t i = 0; i < count; i++)
			out[i] = d(rng64);
	} else if (kind == 3) {
		spice::util::normal_distribution<float> d(static_cast<float>(a), static_cast<float>(b));
		for (std::int64_t i = 0; i < count; i++)
			out[i] = d(rng64);
	} else if (kind == 4) {
		spice::util::binomial_distribution<std::int64_t> d(static_cast<std::int64_t>(a), b);
		for (std::int64_t i = 0; i < count; i++)
			out[i] = static_cast<double>(d(rng64));
	} else if (kind == 5) {
		spice::util::exponential_distribution<float> d(static_cast<float>(a));
		for (std::int64_t i = 0; i < count; i++)
			out[i] = d(rng64);
	} else if (kind == 6) {
		spice::util::seed_seq s{std::seed_seq(il, il + n)}; // by value: only a temporary can be passed (std::seed_seq does not copy)
		std::uint32_t w[8];
		s.generate(reinterpret_cast<std::int32_t*>(w), reinterpret_cast<std::int32_t*>(w) + 8);
		for (std::int64_t i = 0; i < count && i < 8; i++)
			out[i] = static_cast<double>(w[i]);
	}
}

// ---- kahan-compensated dt exactly as snn::step() produces it (snn.cpp:8-10) -------------------
void ref_kahan_dt(float dt, std::int64_t steps, float* out) {
	spice::util::kahan_sum<float> simtime;
	for (std::int64_t i = 0; i < steps; i++) {
		float const d = simtime += dt;
		if (simtime >= 1)
			simtime.reset();
		out[i] = d;
	}
}

// ---- fixed_probability (topology.cpp:75-112) ---------------------------------------------------
std::int64_t ref_fixed_probability_size(std::int64_t src, std::int64_t dst, double p) {
	spice::fixed_probability fp(p);
	fp(src, dst);
	return fp.size();
}

// offsets: src+1 entries, neighbors: >= size() entries.  Returns the edge count.
std::int64_t ref_fixed_probability_generate(std::int64_t src, std::int64_t dst, double p,
                                            std::uint32_t const* il, int n, int increments,
                                            std::int64_t* offsets, std::int32_t* neighbors,
                                            double* seconds_out) {
	spice::fixed_probability fp(p);
	fp(src, dst);
	auto const t0 = clk::now();
	fp.generate(std::span<Int>(reinterpret_cast<Int*>(offsets), static_cast<std::size_t>(src + 1)),
	            std::span<Int32>(neighbors, static_cast<std::size_t>(fp.size())),
	            make_seed(il, n, increments));
	auto const t1 = clk::now();
	if (seconds_out)
		*seconds_out = seconds(t0, t1);
	return (src > 0 && dst > 0 && p > 0) ? offsets[src] : 0;
}

// ---- Brunel (samples/brunel.cpp:78-103), parametrised -----------------------------------------
// Populations in add order P (poisson, N/2), E (lif, 4N/10), I (lif, N/10); six connections in
// the sample's order.  Raster slots are step*3 + {0:P,1:E,2:I}.  state_E/state_I receive the
// final lif::neuron arrays ({float V; int Twait} = 8 bytes each) when non-null.
int ref_brunel_run(std::int64_t N, double p, float w_exc, float w_inh, float dt, float delay,
                   std::uint32_t seed, std::int64_t steps, std::int32_t* ids,
                   std::int64_t capacity, std::int64_t* counts, void* state_E, void* state_I,
                   double* build_seconds, double* sim_seconds, std::int64_t* synaptic_events) {
	using namespace ref_brunel;
	auto const t0 = clk::now();
	spice::snn net(dt, delay, {seed});
	auto P = net.add_population<poisson>(N / 2);
	auto E = net.add_population<lif>(N * 4 / 10);
	auto I = net.add_population<lif>(N / 10);
	net.connect<fixed_weight>(P, E, spice::fixed_probability(p), delay, {w_exc});
	net.connect<fixed_weight>(P, I, spice::fixed_probability(p), delay, {w_exc});
	net.connect<fixed_weight>(E, E, spice::fixed_probability(p), delay, {w_exc});
	net.connect<fixed_weight>(E, I, spice::fixed_probability(p), delay, {w_exc});
	net.connect<fixed_weight>(I, E, spice::fixed_probability(p), delay, {w_inh});
	net.connect<fixed_weight>(I, I, spice::fixed_probability(p), delay, {w_inh});
	auto const t1 = clk::now();

	raster r{ids, capacity, counts};
	double sim = 0;
	for (std::int64_t s = 0; s < steps; s++) {
		auto const a = clk::now();
		net.step();
		sim += seconds(a, clk::now());
		if (counts) {
			r.push(P->spikes(0), s * 3 + 0);
			r.push(E->spikes(0), s * 3 + 1);
			r.push(I->spikes(0), s * 3 + 2);
		}
	}
	if (state_E)
		std::memcpy(state_E, E->get_neurons().data(), E->size() * sizeof(lif::neuron));
	if (state_I)
		std::memcpy(state_I, I->get_neurons().data(), I->size() * sizeof(lif::neuron));
	if (build_seconds)
		*build_seconds = seconds(t0, t1);
	if (sim_seconds)
		*sim_seconds = sim;
	// Tally of Syn::deliver invocations (synapse_population.h:118-133), recomputed outside the
	// timed region: regenerate each connection's offsets from its seed (the stateless P burns no
	// seed, E and I one each, so connection j was built from the (2 + j)-th increment) and add the
	// out-degree of every spike that was delivered, i.e. emitted at a step <= steps - delay.
	if (synaptic_events && counts && !r.overflow) {
		std::int64_t const d   = static_cast<std::int64_t>(std::round(delay / dt));
		std::int64_t const nsz[3] = {N / 2, N * 4 / 10, N / 10};
		int const csrc[6] = {0, 0, 1, 1, 2, 2}, cdst[6] = {1, 2, 1, 2, 1, 2};
		std::int64_t total = 0;
		for (int j = 0; j < 6; j++) {
			spice::fixed_probability fp(p);
			fp(nsz[csrc[j]], nsz[cdst[j]]);
			std::vector<Int> off(static_cast<std::size_t>(nsz[csrc[j]]) + 1);
			std::vector<Int32> nb(static_cast<std::size_t>(fp.size()));
			std::uint32_t il[1] = {seed};
			fp.generate(off, nb, make_seed(il, 1, 2 + j));
			std::int64_t at = 0;
			for (std::int64_t s = 0; s < steps; s++)
				for (int pop = 0; pop < 3; pop++) {
					std::int64_t const c = counts[s * 3 + pop];
					if (pop == csrc[j] && s <= steps - d)
						for (std::int64_t k = 0; k < c; k++)
							total += off[ids[at + k] + 1] - off[ids[at + k]];
					at += c;
				}
		}
		*synaptic_events = total;
	}
	return r.overflow ? 1 : 0;
}

// ---- Brunel, incremental: build once, advance in slices (bench.py's reference arm) --------------
// The timed region is net.step() only; the tally of Syn::deliver invocations (out-degree of every
// spike delivered in the slice: a spike of age delay-1 is delivered at the end of a step,
// snn.cpp:21-25) is recomputed outside it from regenerated offsets.
struct ref_brunel_net {
	spice::snn net;
	spice::detail::neuron_population<ref_brunel::poisson>* P;
	spice::detail::neuron_population<ref_brunel::lif>* E;
	spice::detail::neuron_population<ref_brunel::lif>* I;
	std::vector<Int> outdeg[3]; // per source population: summed out-degree over its two connections
	std::int64_t d, steps_run = 0;
	ref_brunel_net(float dt, float delay, std::uint32_t seed) : net(dt, delay, {seed}) {}
};

void* ref_brunel_open(std::int64_t N, double p, float w_exc, float w_inh, float dt, float delay, std::uint32_t seed,
                      double* build_seconds) {
	using namespace ref_brunel;
	auto const t0 = clk::now();
	auto* h = new ref_brunel_net(dt, delay, seed);
	h->P = h->net.add_population<poisson>(N / 2);
	h->E = h->net.add_population<lif>(N * 4 / 10);
	h->I = h->net.add_population<lif>(N / 10);
	h->net.connect<fixed_weight>(h->P, h->E, spice::fixed_probability(p), delay, {w_exc});
	h->net.connect<fixed_weight>(h->P, h->I, spice::fixed_probability(p), delay, {w_exc});
	h->net.connect<fixed_weight>(h->E, h->E, spice::fixed_probability(p), delay, {w_exc});
	h->net.connect<fixed_weight>(h->E, h->I, spice::fixed_probability(p), delay, {w_exc});
	h->net.connect<fixed_weight>(h->I, h->E, spice::fixed_probability(p), delay, {w_inh});
	h->net.connect<fixed_weight>(h->I, h->I, spice::fixed_probability(p), delay, {w_inh});
	if (build_seconds)
		*build_seconds = seconds(t0, clk::now());
	h->d = static_cast<std::int64_t>(std::round(delay / dt));
	std::int64_t const nsz[3] = {N / 2, N * 4 / 10, N / 10};
	int const csrc[6] = {0, 0, 1, 1, 2, 2}, cdst[6] = {1, 2, 1, 2, 1, 2};
	for (int s = 0; s < 3; s++)
		h->outdeg[s].assign(static_cast<std::size_t>(nsz[s]), 0);
	for (int j = 0; j < 6; j++) {
		spice::fixed_probability fp(p);
		fp(nsz[csrc[j]], nsz[cdst[j]]);
		std::vector<Int> off(static_cast<std::size_t>(nsz[csrc[j]]) + 1);
		std::vector<Int32> nb(static_cast<std::size_t>(fp.size()));
		std::uint32_t il[1] = {seed};
		fp.generate(off, nb, make_seed(il, 1, 2 + j));
		for (std::int64_t i = 0; i < nsz[csrc[j]]; i++)
			h->outdeg[csrc[j]][static_cast<std::size_t>(i)] += off[i + 1] - off[i];
	}
	return h;
}

int ref_brunel_advance(void* handle, std::int64_t steps, double* sim_seconds, std::int64_t* synaptic_events,
                       std::int64_t* spikes) {
	auto* h = static_cast<ref_brunel_net*>(handle);
	double sim = 0;
	std::int64_t ev = 0, sp = 0;
	for (std::int64_t s = 0; s < steps; s++) {
		auto const a = clk::now();
		h->net.step();
		sim += seconds(a, clk::now());
		h->steps_run++;
		sp += static_cast<std::int64_t>(h->P->spikes(0).size() + h->E->spikes(0).size() + h->I->spikes(0).size());
		if (h->steps_run >= h->d) {
			for (auto id : h->P->spikes(h->d - 1)) ev += h->outdeg[0][id];
			for (auto id : h->E->spikes(h->d - 1)) ev += h->outdeg[1][id];
			for (auto id : h->I->spikes(h->d - 1)) ev += h->outdeg[2][id];
		}
	}
	if (sim_seconds) *sim_seconds = sim;
	if (synaptic_events) *synaptic_events = ev;
	if (spikes) *spikes = sp;
	return 0;
}

void ref_brunel_close(void* handle) { delete static_cast<ref_brunel_net*>(handle); }

// ---- SSSP (samples/sssp.cpp:100-116): the reference's own vertex / edge models on its own graph ----
// distances[7] after vertices - 1 steps (DeliverFromTo synapses, per-synapse and per-population init)
int ref_sssp_distances(std::int64_t* distances) {
	using namespace ref_sssp;
	spice::snn sssp(1, 1, {1337});
	auto vertices = sssp.add_population<vertex>(7, {0});
	spice::adj_list adj;
	for (Int src : range(7))
		for (Int dst : range(7))
			if (adj_matrix[src][dst])
				adj.connect(src, dst);
	sssp.connect<edge>(vertices, vertices, adj, 1);
	for (Int i : range(vertices->size() - 1)) {
		sssp.step();
		(void)i;
	}
	for (int v = 0; v < 7; v++)
		distances[v] = vertices->get_neurons()[v].distance;
	return 0;
}

// ---- Brunel+ (samples/brunel+.cpp:102-117): E->E plastic -------------------------------------
int ref_brunel_plus_run(std::int64_t N, double p, float w_exc, float w_inh, float dt, float delay,
                        std::uint32_t seed, std::int64_t steps, std::int32_t* ids,
                        std::int64_t capacity, std::int64_t* counts, void* state_E, void* state_I,
                        double* build_seconds, double* sim_seconds) {
	using namespace ref_brunel_plus;
	auto const t0 = clk::now();
	spice::snn net(dt, delay, {seed});
	auto P = net.add_population<poisson>(N / 2);
	auto E = net.add_population<lif>(N * 4 / 10);
	auto I = net.add_population<lif>(N / 10);
	net.connect<fixed_weight>(P, E, spice::fixed_probability(p), delay, {w_exc});
	net.connect<fixed_weight>(P, I, spice::fixed_probability(p), delay, {w_exc});
	net.connect<plastic>(E, E, spice::fixed_probability(p), delay);
	net.connect<fixed_weight>(E, I, spice::fixed_probability(p), delay, {w_exc});
	net.connect<fixed_weight>(I, E, spice::fixed_probability(p), delay, {w_inh});
	net.connect<fixed_weight>(I, I, spice::fixed_probability(p), delay, {w_inh});
	auto const t1 = clk::now();

	raster r{ids, capacity, counts};
	double sim = 0;
	for (std::int64_t s = 0; s < steps; s++) {
		auto const a = clk::now();
		net.step();
		sim += seconds(a, clk::now());
		if (counts) {
			r.push(P->spikes(0), s * 3 + 0);
			r.push(E->spikes(0), s * 3 + 1);
			r.push(I->spikes(0), s * 3 + 2);
		}
	}
	if (state_E)
		std::memcpy(state_E, E->get_neurons().data(), E->size() * sizeof(lif::neuron));
	if (state_I)
		std::memcpy(state_I, I->get_neurons().data(), I->size() * sizeof(lif::neuron));
	if (build_seconds)
		*build_seconds = seconds(t0, t1);
	if (sim_seconds)
		*sim_seconds = sim;
	return r.overflow ? 1 : 0;
}

// ---- Vogels (samples/vogels.cpp:62-76): E (8N/10), I (2N/10), static synapses ------------------
// Raster slots are step*2 + {0:E,1:I}.  lif::neuron = {float V, Gex, Gin; int32 Twait} = 16 B.
int ref_vogels_run(std::int64_t N, double p, float w_exc, float w_inh, float dt, float delay,
                   std::uint32_t seed, std::int64_t steps, std::int32_t* ids,
                   std::int64_t capacity, std::int64_t* counts, void* state_E, void* state_I,
                   double* build_seconds, double* sim_seconds) {
	using namespace ref_vogels;
	auto const t0 = clk::now();
	spice::snn net(dt, delay, {seed});
	auto E = net.add_population<lif>(N * 8 / 10);
	auto I = net.add_population<lif>(N * 2 / 10);
	net.connect<excitatory>(E, E, spice::fixed_probability(p), delay, {w_exc});
	net.connect<excitatory>(E, I, spice::fixed_probability(p), delay, {w_exc});
	net.connect<inhibitory>(I, E, spice::fixed_probability(p), delay, {w_inh});
	net.connect<inhibitory>(I, I, spice::fixed_probability(p), delay, {w_inh});
	auto const t1 = clk::now();

	raster r{ids, capacity, counts};
	double sim = 0;
	for (std::int64_t s = 0; s < steps; s++) {
		auto const a = clk::now();
		net.step();
		sim += seconds(a, clk::now());
		if (counts) {
			r.push(E->spikes(0), s * 2 + 0);
			r.push(I->spikes(0), s * 2 + 1);
		}
	}
	if (state_E)
		std::memcpy(state_E, E->get_neurons().data(), E->size() * sizeof(lif::neuron));
	if (state_I)
		std::memcpy(state_I, I->get_neurons().data(), I->size() * sizeof(lif::neuron));
	if (build_seconds)
		*build_seconds = seconds(t0, t1);
	if (sim_seconds)
		*sim_seconds = sim;
	return r.overflow ? 1 : 0;
}

// libm values at the reference's call sites (random.h:271 -> glibc log; brunel+.cpp:78-79,96-97
// -> glibc expf / pow), so tests can pin the device restatements against this host's libm.
// Called through volatile pointers so -ffast-math cannot swap in a libmvec vector variant:
// the reference's call sites are scalar calls inside serial loops.
void ref_libm_log(double const* x, std::int64_t n, double* out) {
	double (*volatile f)(double) = static_cast<double (*)(double)>(&std::log);
	for (std::int64_t i = 0; i < n; i++)
		out[i] = f(x[i]);
}
void ref_libm_expf(float const* x, std::int64_t n, float* out) {
	float (*volatile f)(float) = static_cast<float (*)(float)>(&std::exp);
	for (std::int64_t i = 0; i < n; i++)
		out[i] = f(x[i]);
}
void ref_libm_pow(double const* x, double const* y, std::int64_t n, double* out) {
	double (*volatile f)(double, double) = static_cast<double (*)(double, double)>(&std::pow);
	for (std::int64_t i = 0; i < n; i++)
		out[i] = f(x[i], y[i]);
}

char const* ref_build_flavour() {
#ifdef REF_STRICT
	return "strict (-O2 -fno-fast-math -ffp-contract=off)";
#else
	return "reference flags (-O2 -ffast-math ... -march=haswell -mfpmath=sse)";
#endif
}
}
