// The reference's unit tests of the hot path (test/detail/synapse_population.cpp:28-81,
// test/detail/neuron_population.cpp:146-170) exercise internal classes; here the same small graphs
// and fake models run through the PUBLIC API on the GPU: a host-fed source population
// (per-population update()) emits the test's spike list, an adj_list carries the 3 x 5 graph, and the
// counters the synapses leave in the target neurons are read back with get_neurons().
// Exit status 0 = every expectation of the reference tests holds.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "spice/snn.h"

using namespace spice;
using namespace spice::util;

#define EXPECT_EQ(a, b)                                                                                          \
	do {                                                                                                         \
		if (!((a) == (b))) {                                                                                     \
			std::printf("%s:%d: expected %s == %s (%lld vs %lld)\n", __FILE__, __LINE__, #a, #b, (long long)(a), \
			            (long long)(b));                                                                         \
			std::exit(1);                                                                                        \
		}                                                                                                        \
	} while (0)

// emits {0, 1} in the first step, nothing afterwards (synapse_population.cpp:47 `spikes[] = {0, 1}`)
struct source {
	Int step = 0;
	void update(float, auto&, std::vector<Int32>& out) {
		if (step++ == 0)
			out.insert(out.end(), {0, 1});
	}
};
static_assert(PerPopulationUpdate<source>);

// emits {1, 2} in step `at` (synapse_population.cpp:100 `spikes[] = {1, 2}`)
struct late_source {
	Int at   = 0;
	Int step = 0;
	void update(float, auto&, std::vector<Int32>& out) {
		if (step++ == at)
			out.insert(out.end(), {1, 2});
	}
};

struct stateful_neuron { // synapse_population.cpp:13-19
	struct neuron {
		int received_count = 0;
	};
	SPICE_HD bool update(neuron&, float, auto&) const { return false; }
};
static_assert(StatefulNeuron<stateful_neuron>);

struct stateless_synapse { // synapse_population.cpp:22-24
	SPICE_HD void deliver(stateful_neuron::neuron& n) const { n.received_count++; }
};
struct stateful_synapse { // synapse_population.cpp:59-66
	struct synapse {
		int w = 2;
	};
	SPICE_HD void deliver(synapse const& syn, stateful_neuron::neuron& n) const { n.received_count += syn.w; }
};

struct plastic_synapse { // synapse_population.cpp:83-93: counts how many steps it was brought forward
	struct synapse {
		int update_count = 0;
	};
	SPICE_HD void deliver(synapse const& syn, stateful_neuron::neuron& n) const { n.received_count = syn.update_count; }
	SPICE_HD void update(synapse& syn, float, bool, bool) const { syn.update_count++; }
	SPICE_HD void skip(synapse& syn, float, Int steps) const { syn.update_count += steps; }
};
static_assert(PlasticSynapse<plastic_synapse>);

// concepts.h:76-99 DeliverFromTo without synapse state (synapse_population.h:125-131): a source that fires once and
// carries a value its synapses read
struct tagged_source {
	struct neuron {
		int tag   = 0;
		int fired = 0;
	};
	SPICE_HD void init(neuron& n, Int id, auto&) const { n.tag = 10 * (static_cast<int>(id) + 1); }
	SPICE_HD bool update(neuron& n, float, auto&) const {
		if (n.fired)
			return false;
		n.fired = 1;
		return true;
	}
};
struct from_to_stateless {
	SPICE_HD void deliver(tagged_source::neuron const& from, stateful_neuron::neuron& to) const { to.received_count += from.tag; }
};
static_assert(Synapse<from_to_stateless, tagged_source, stateful_neuron> && !StatefulSynapse<from_to_stateless>);

// a stateless neuron that declares one draw per update() and takes two: every neuron behind it in its chunk of the step's
// stream would read the wrong draws, so the run must fail loudly (the rng_draws contract, DESIGN.md boundary)
struct greedy {
	static constexpr int rng_draws = 1;
	SPICE_HD bool update(float, auto& rng) const { return ((rng() ^ rng()) & 7) == 0; }
};
static_assert(StatelessNeuron<greedy>);
struct greedy_ok { // the same with its one draw declared
	static constexpr int rng_draws = 1;
	SPICE_HD bool update(float, auto& rng) const { return (rng() & 7) == 0; }
};

// a STATEFUL neuron that draws (the everyday "LIF with noise"): neuron_population.h:116-124 updates the neurons in index
// order with the step's one engine, so neuron i of a population takes draws [offset + 2 i, offset + 2 i + 2) of the stream
struct noisy {
	static constexpr int rng_draws = 2;
	struct neuron {
		float v   = 0;
		int fired = 0;
	};
	SPICE_HD bool update(neuron& n, float, auto& rng) const {
		uniform_real_distribution<float> kick(0.0f, 0.3f);
		n.v += kick(rng);
		if ((rng() & 7) == 0) // a second draw that sometimes takes the charge away again
			n.v *= 0.5f;
		if (n.v < 1.0f)
			return false;
		n.v = 0;
		n.fired++;
		return true;
	}
};
static_assert(StatefulNeuron<noisy>);

// a per-population update() that draws (concepts.h:46-57): two neighbouring neurons fire, chosen by the step's engine
struct roulette {
	Int n = 0;
	void update(float, auto& rng, std::vector<Int32>& out) {
		Int const first = static_cast<Int>(rng() % static_cast<UInt>(n));
		out.push_back(static_cast<Int32>(first));
		if (rng() & 1)
			out.push_back(static_cast<Int32>((first + 1) % n));
	}
};
static_assert(PerPopulationUpdate<roulette>);

adj_list graph() { // synapse_population.cpp:33-40
	adj_list adj;
	adj.connect(0, 0);
	adj.connect(0, 1);
	adj.connect(0, 3);
	adj.connect(1, 3);
	adj.connect(2, 4);
	return adj;
}

// The same 3 x 5 graph as a user-defined Topology writing an edge_stream (topology.h:11-35, topology.cpp:12-55):
// the reference's extension point.  It also checks the seed it is handed: the connection's own (synapse_population.h:31).
struct graph_topology : Topology {
	UInt128 seen{};
	Int size() const override { return 8; } // an upper bound, like fixed_probability::size()
	using Topology::generate;
	void generate(edge_stream& s, util::seed_seq const& seed) override {
		seen = seed.seed();
		s << std::pair{Int32(0), Int32(0)} << std::pair{Int32(0), Int32(1)} << std::pair{Int32(0), Int32(3)} << std::pair{Int32(1), Int32(3)}
		  << std::pair{Int32(2), Int32(4)};
	}
};

template <class Syn>
std::vector<int> deliver_once() {
	snn net(1, 1, {1337});
	auto src = net.add_population<source>(3);
	auto dst = net.add_population<stateful_neuron>(5);
	auto adj = graph();
	net.connect<Syn>(src, dst, adj, 1);
	net.step();
	std::vector<int> got;
	for (auto const& n : dst->get_neurons())
		got.push_back(n.received_count);
	return got;
}

// neuron_population.cpp:141-170
struct per_population_update {
	void update(float, auto&, std::vector<Int32>& spikes) { spikes.insert(spikes.end(), {1, 3, 8}); }
};

int main() {
	{ // SynapsePopulation.DeliverStateless
		auto const n = deliver_once<stateless_synapse>();
		EXPECT_EQ(n[0], 1);
		EXPECT_EQ(n[1], 1);
		EXPECT_EQ(n[2], 0);
		EXPECT_EQ(n[3], 2);
		EXPECT_EQ(n[4], 0);
	}
	{ // SynapsePopulation.DeliverStateful
		auto const n = deliver_once<stateful_synapse>();
		EXPECT_EQ(n[0], 2);
		EXPECT_EQ(n[1], 2);
		EXPECT_EQ(n[2], 0);
		EXPECT_EQ(n[3], 4);
		EXPECT_EQ(n[4], 0);
	}
	{ // a stateless synapse that reads its source (no reference test; semantics of synapse_population.h:125-131)
		snn net(1, 1, {1337});
		auto src = net.add_population<tagged_source>(3);
		auto dst = net.add_population<stateful_neuron>(5);
		auto adj = graph();
		net.connect<from_to_stateless>(src, dst, adj, 1);
		net.step();
		net.step(); // nobody fires twice
		auto const n = dst->get_neurons();
		EXPECT_EQ(n[0].received_count, 10);
		EXPECT_EQ(n[1].received_count, 10);
		EXPECT_EQ(n[2].received_count, 0);
		EXPECT_EQ(n[3].received_count, 30);
		EXPECT_EQ(n[4].received_count, 30);
	}
	// SynapsePopulation.DeliverPlastic (synapse_population.cpp:96-168): the lazy bookkeeping (`_ages`, the flush of
	// snn.cpp:17-19 at step 0).  A synapse delivered at step T has been brought forward over steps 0..T exactly
	// once each, however often it was visited: T + 1 updates (the reference's cases T = 0, 1 and 9).
	for (Int T : {0, 1, 9, 70}) {
		snn net(1, 1, {1337});
		auto src = net.add_population<late_source>(3, {T});
		auto dst = net.add_population<stateful_neuron>(5);
		auto adj = graph();
		net.connect<plastic_synapse>(src, dst, adj, 1);
		for (Int i = 0; i <= T; i++)
			net.step();
		auto const n = dst->get_neurons();
		EXPECT_EQ(n[3].received_count, T + 1);
		EXPECT_EQ(n[4].received_count, T + 1);
		EXPECT_EQ(n[0].received_count, 0);
	}
	{ // NeuronPopulation.PerPopulationUpdate
		snn net(1, 1, {1337});
		auto pop = net.add_population<per_population_update>(10);
		EXPECT_EQ(pop->size(), 10);
		net.step();
		EXPECT_EQ(pop->spikes(0).size(), 3u);
		EXPECT_EQ(pop->spikes(0)[0], 1);
		EXPECT_EQ(pop->spikes(0)[1], 3);
		EXPECT_EQ(pop->spikes(0)[2], 8);
	}
	{ // a user-defined Topology (edge_stream) delivers like the adj_list of the same edges, stateless and stateful
		snn net(1, 1, {1337});
		auto src  = net.add_population<source>(3);          // per-population update: no seed++
		auto dst  = net.add_population<stateful_neuron>(5); // stateful adapter: one seed++ (neuron_population.h:60-67)
		auto dst2 = net.add_population<stateful_neuron>(5);
		graph_topology topo;
		net.connect<stateless_synapse>(src, dst, topo, 1);
		util::seed_seq want{1337};
		want++;
		want++;
		EXPECT_EQ(topo.seen.lo, want.seed().lo);
		EXPECT_EQ(topo.seen.hi, want.seed().hi);
		net.connect<stateful_synapse>(src, dst2, graph_topology{}, 1);
		net.step();
		auto const a = dst->get_neurons();
		auto const b = dst2->get_neurons();
		int const expect[5] = {1, 1, 0, 2, 0};
		for (int i = 0; i < 5; i++) {
			EXPECT_EQ(a[i].received_count, expect[i]);
			EXPECT_EQ(b[i].received_count, 2 * expect[i]);
		}
	}
	{ // preconditions throw std::logic_error (util/assert.h:3-17): a delay beyond max_delay (snn.h:36-38)
		snn net(1, 1, {1337});
		auto src = net.add_population<source>(3);
		auto dst = net.add_population<stateful_neuron>(5);
		auto adj = graph();
		bool threw = false;
		try {
			net.connect<stateless_synapse>(src, dst, adj, 2);
		} catch (std::logic_error const&) {
			threw = true;
		}
		EXPECT_EQ(threw, true);
	}
	{ // stateful neurons that draw, two populations of them behind a stateless one that draws as well: every population's
	  // draws start where the population before it stopped (snn.cpp:12-15), checked against the loop run on the host
		snn net(1, 1, {4711});
		Int const n0 = 77, n1 = 333, n2 = 150;
		auto pre = net.add_population<greedy_ok>(n0);
		auto a   = net.add_population<noisy>(n1);
		auto b   = net.add_population<noisy>(n2);
		auto sink = net.add_population<stateful_neuron>(5); // behind the drawing populations: no draws, one seed++
		net.connect<stateless_synapse>(pre, sink, fixed_probability(0.5), 1);
		uint64_t sd[2];
		EXPECT_EQ(spice_ctx_seed(net.context(), sd), SPICE_OK);
		util::seed_seq seed(UInt128{sd[0], sd[1]});
		std::vector<noisy::neuron> ha(n1), hb(n2);
		for (int step = 0; step < 40; step++) {
			util::xoroshiro64_128p rng(seed++);
			std::vector<Int32> s0, sa, sb;
			greedy_ok g;
			noisy m;
			for (Int i = 0; i < n0; i++)
				if (g.update(1.0f, rng))
					s0.push_back(static_cast<Int32>(i));
			for (Int i = 0; i < n1; i++)
				if (m.update(ha[i], 1.0f, rng))
					sa.push_back(static_cast<Int32>(i));
			for (Int i = 0; i < n2; i++)
				if (m.update(hb[i], 1.0f, rng))
					sb.push_back(static_cast<Int32>(i));
			net.step();
			auto check = [](std::span<Int32 const> got, std::vector<Int32> const& want) {
				EXPECT_EQ(got.size(), want.size());
				for (std::size_t k = 0; k < want.size(); k++)
					EXPECT_EQ(got[k], want[k]);
			};
			check(pre->spikes(0), s0);
			check(a->spikes(0), sa);
			check(b->spikes(0), sb);
		}
		auto ga = a->get_neurons();
		auto gb = b->get_neurons();
		long long fired = 0;
		for (Int i = 0; i < n1; i++) {
			EXPECT_EQ(ga[i].v == ha[i].v, true);
			EXPECT_EQ(ga[i].fired, ha[i].fired);
			fired += ha[i].fired;
		}
		for (Int i = 0; i < n2; i++)
			EXPECT_EQ(gb[i].v == hb[i].v && gb[i].fired == hb[i].fired, true);
		EXPECT_EQ(fired > 100, true);
	}
	{ // a host-fed population that draws, behind a device population that draws: it continues the step's stream where the
	  // device population stopped (snn.cpp:12-15), and its own draws (1 or 2 per step) move nobody else
		snn net(1, 1, {99});
		Int const n0 = 45, n1 = 20;
		auto pre  = net.add_population<greedy_ok>(n0);
		auto host = net.add_population<roulette>(n1, roulette{n1});
		auto sink = net.add_population<stateful_neuron>(n1);
		net.connect<stateless_synapse>(host, sink, fixed_probability(1.0), 1);
		uint64_t sd[2];
		EXPECT_EQ(spice_ctx_seed(net.context(), sd), SPICE_OK);
		util::seed_seq seed(UInt128{sd[0], sd[1]});
		long long delivered = 0;
		for (int step = 0; step < 30; step++) {
			util::xoroshiro64_128p rng(seed++);
			std::vector<Int32> s0, s1;
			greedy_ok g;
			for (Int i = 0; i < n0; i++)
				if (g.update(1.0f, rng))
					s0.push_back(static_cast<Int32>(i));
			roulette{n1}.update(1.0f, rng, s1);
			delivered += static_cast<long long>(s1.size()) * n1;
			net.step();
			auto a = pre->spikes(0), b = host->spikes(0);
			EXPECT_EQ(a.size(), s0.size());
			for (std::size_t k = 0; k < s0.size(); k++)
				EXPECT_EQ(a[k], s0[k]);
			std::sort(s1.begin(), s1.end()); // spikes() of a host-fed population comes back ascending
			EXPECT_EQ(b.size(), s1.size());
			for (std::size_t k = 0; k < s1.size(); k++)
				EXPECT_EQ(b[k], s1[k]);
		}
		long long got = 0;
		for (auto const& n : sink->get_neurons())
			got += n.received_count;
		EXPECT_EQ(got, delivered); // delay 1: delivered at the end of the step that emitted them (get_neurons folds them in)
	}
	{ // get_neurons() is a read-only copy, set_neurons() carries changes back (the state lives on the device)
		snn net(1, 1, {1337});
		auto src = net.add_population<source>(3);
		auto dst = net.add_population<stateful_neuron>(5);
		auto adj = graph();
		net.connect<stateless_synapse>(src, dst, adj, 1);
		net.step(); // received_count = {1, 1, 0, 2, 0}
		auto const view = dst->get_neurons();
		std::vector<stateful_neuron::neuron> mine(view.begin(), view.end());
		for (auto& n : mine)
			n.received_count += 100;
		dst->set_neurons(mine);
		net.step(); // nothing fires any more: the pending deliveries were part of the copy and are not applied twice
		int const expect[5] = {101, 101, 100, 102, 100};
		auto const after = dst->get_neurons();
		for (int i = 0; i < 5; i++)
			EXPECT_EQ(after[i].received_count, expect[i]);
	}
	{ // rng_draws is checked in both directions: drawing more than declared is reported by the next synchronising call
		snn net(1, 1, {1337});
		auto src = net.add_population<greedy>(100);
		auto dst = net.add_population<stateful_neuron>(5);
		net.connect<stateless_synapse>(src, dst, fixed_probability(0.5), 1);
		bool threw = false;
		try {
			net.step();
			(void)src->spikes(0);
		} catch (std::logic_error const&) {
			threw = true;
		}
		EXPECT_EQ(threw, true);
	}
	std::printf("facade_spec: ok\n");
	return 0;
}
