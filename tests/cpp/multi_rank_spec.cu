// More than one rank for what the reference's samples use beside fixed_probability + stateless synapses: per-synapse init
// hooks (synapse_population.h:34-41) over fixed_probability and adj_list connections, host-fed populations
// (per-population update(), neuron_population.h:86-101) and synapses whose deliver() reads the source neuron
// (DeliverFromTo, concepts.h:76-99).  Two rank contexts of one process on one device (peer handles
// resolved without CUDA IPC) must reproduce the one-rank run of this backend — which the samples and facade_spec pin to the
// reference — spike for spike, neuron for neuron and synapse for synapse.
// Exit status 0 = every expectation holds.
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include "spice/snn.h"

using namespace spice;
using namespace spice::util;

#define EXPECT(c)                                                       \
	do {                                                                \
		if (!(c)) {                                                     \
			std::printf("%s:%d: expected %s\n", __FILE__, __LINE__, #c); \
			std::exit(1);                                               \
		}                                                               \
	} while (0)

// SPICE_SPEC_TWO_DEVICES=1 (and two GPUs): rank r lives on device r, so every peer store crosses NVLink
static int device_of(int rank) {
	static int const two = [] {
		int n = 0;
		char const* e = std::getenv("SPICE_SPEC_TWO_DEVICES");
		return (e && e[0] == '1' && cudaGetDeviceCount(&n) == cudaSuccess && n >= 2) ? 1 : 0;
	}();
	return two ? rank : 0;
}

// a deterministic spike train: neuron i fires in step t when (i * 7 + t * 13) % 17 == 0
struct drummer {
	Int n    = 0;
	Int step = 0;
	void update(float, auto&, std::vector<Int32>& out) {
		for (Int i = 0; i < n; i++)
			if ((i * 7 + step * 13) % 17 == 0)
				out.push_back(static_cast<Int32>(i));
		step++;
	}
};
static_assert(PerPopulationUpdate<drummer>);

struct integrator {
	struct neuron {
		float v   = 0;
		int count = 0;
	};
	SPICE_HD bool update(neuron& n, float, auto&) const {
		bool const fire = n.v >= 1.0f;
		if (fire)
			n.v = 0;
		return fire;
	}
};
static_assert(StatefulNeuron<integrator>);

// per-synapse init hook: the weight depends on (src, dst) and on the hook's engine, i.e. on every synapse before it
struct drawn_weight {
	struct synapse {
		float w = 0;
	};
	SPICE_HD void init(synapse& syn, Int src, Int dst, auto& rng) const {
		uniform_real_distribution<float> u(0.05f, 0.3f);
		syn.w = u(rng) + 1e-3f * static_cast<float>((src + 3 * dst) % 5);
	}
	SPICE_HD void deliver(synapse const& syn, integrator::neuron& n) const {
		n.v += syn.w;
		n.count++;
	}
};
static_assert(StatefulSynapse<drawn_weight>);

// DeliverFromTo (concepts.h:76-99, synapse_population.h:125-131): deliver() reads the SOURCE neuron as it is at the end of the
// step in which it fired — on another rank, for most synapses of a sharded network
struct charger {
	struct neuron {
		float charge = 0;
		int phase    = 0;
	};
	SPICE_HD void init(neuron& n, Int id, auto&) const { n.phase = static_cast<int>(id % 11); }
	SPICE_HD bool update(neuron& n, float, auto&) const {
		n.charge += 0.125f * static_cast<float>(1 + n.phase % 3);
		n.phase++;
		if (n.phase % 11 != 0)
			return false;
		n.charge *= 0.5f;
		return true;
	}
};
static_assert(StatefulNeuron<charger>);
struct carry_charge {
	struct synapse {
		float gain = 1;
	};
	SPICE_HD void init(synapse& syn, Int src, Int dst, auto& rng) const {
		uniform_real_distribution<float> u(0.5f, 1.5f);
		syn.gain = u(rng) + 0.01f * static_cast<float>((src + dst) % 3);
	}
	SPICE_HD void deliver(synapse const& syn, charger::neuron const& from, integrator::neuron& to) const {
		to.v += 0.01f * syn.gain * from.charge;
		to.count++;
	}
};
static_assert(StatefulSynapse<carry_charge>);

struct from_to_network {
	std::unique_ptr<snn> net;
	spice::detail::neuron_population<charger>* C    = nullptr;
	spice::detail::neuron_population<integrator>* T = nullptr;
};

static from_to_network build_from_to(int rank, int world) {
	from_to_network w;
	w.net = std::make_unique<snn>(1e-3f, 3e-3f, seed_seq{21}, device_of(rank), rank, world);
	w.C   = w.net->add_population<charger>(407);
	w.T   = w.net->add_population<integrator>(311);
	w.net->connect<carry_charge>(w.C, w.T, fixed_probability(0.08), 2e-3f);
	w.net->connect<carry_charge>(w.C, w.T, fixed_probability(0.04), 3e-3f);
	return w;
}

static void connect_ranks(snn* a, snn* b) {
	std::vector<unsigned char> handles;
	int64_t each = 0;
	for (snn* n : {a, b}) {
		EXPECT(spice_ctx_finalize(n->context()) == SPICE_OK);
		int64_t bytes = 0;
		EXPECT(spice_ctx_peer_handle(n->context(), nullptr, &bytes) == SPICE_OK);
		each = bytes;
		handles.resize(handles.size() + static_cast<size_t>(bytes));
		EXPECT(spice_ctx_peer_handle(n->context(), handles.data() + handles.size() - bytes, &bytes) == SPICE_OK);
	}
	for (snn* n : {a, b})
		EXPECT(spice_ctx_set_peers(n->context(), handles.data(), each) == SPICE_OK);
}

static void from_to_on_two_ranks() {
	from_to_network one = build_from_to(0, 1);
	from_to_network r[2] = {build_from_to(0, 2), build_from_to(1, 2)};
	connect_ranks(r[0].net.get(), r[1].net.get());
	long long spikes = 0;
	for (int step = 0; step < 60; step++) {
		one.net->step();
		for (auto& w : r)
			w.net->step();
		for (auto& w : r)
			for (Int p = 0; p < 2; p++) {
				auto got = w.net->spikes(p), want = one.net->spikes(p);
				EXPECT(got.size() == want.size());
				for (std::size_t i = 0; i < got.size(); i++)
					EXPECT(got[i] == want[i]);
			}
		spikes += static_cast<long long>(one.net->spikes(0).size() + one.net->spikes(1).size());
	}
	auto t1 = one.T->get_neurons();
	std::vector<integrator::neuron> t2;
	for (auto& w : r) {
		auto t = w.T->get_neurons();
		t2.insert(t2.end(), t.begin(), t.end());
	}
	EXPECT(t1.size() == t2.size());
	long long delivered = 0;
	for (std::size_t i = 0; i < t1.size(); i++) {
		EXPECT(t1[i].v == t2[i].v && t1[i].count == t2[i].count);
		delivered += t1[i].count;
	}
	EXPECT(spikes > 1000 && delivered > 10000);
	std::printf("from_to on two ranks ok: %lld spikes, %lld deliveries\n", spikes, delivered);
}

struct network {
	std::unique_ptr<snn> net;
	spice::detail::neuron_population<drummer>* D    = nullptr;
	spice::detail::neuron_population<integrator>* A = nullptr;
	spice::detail::neuron_population<integrator>* B = nullptr;
};

static network build(int rank, int world) {
	network w;
	w.net = std::make_unique<snn>(1e-3f, 4e-3f, seed_seq{7, 11}, device_of(rank), rank, world);
	w.D   = w.net->add_population<drummer>(300, drummer{300});
	w.A   = w.net->add_population<integrator>(501);
	w.B   = w.net->add_population<integrator>(233);
	w.net->connect<drawn_weight>(w.D, w.A, fixed_probability(0.05), 2e-3f);
	adj_list adj;
	for (Int i = 0; i < 501; i++)
		for (Int k = 1; k <= 1 + i % 4; k++)
			adj.connect(static_cast<Int32>(i), static_cast<Int32>((i * k + 17 * k) % 233));
	w.net->connect<drawn_weight>(w.A, w.B, adj, 3e-3f);
	w.net->connect<drawn_weight>(w.B, w.A, fixed_probability(0.03), 4e-3f);
	return w;
}

template <class P>
static std::vector<integrator::neuron> state(P* pop) {
	auto span = pop->get_neurons();
	return {span.begin(), span.end()};
}

int main() {
	network one = build(0, 1);
	network r[2] = {build(0, 2), build(1, 2)};
	std::vector<unsigned char> handles;
	int64_t each = 0;
	for (auto& w : r) {
		EXPECT(spice_ctx_finalize(w.net->context()) == SPICE_OK);
		int64_t n = 0;
		EXPECT(spice_ctx_peer_handle(w.net->context(), nullptr, &n) == SPICE_OK);
		each = n;
		handles.resize(handles.size() + static_cast<size_t>(n));
		EXPECT(spice_ctx_peer_handle(w.net->context(), handles.data() + handles.size() - n, &n) == SPICE_OK);
	}
	for (auto& w : r)
		EXPECT(spice_ctx_set_peers(w.net->context(), handles.data(), each) == SPICE_OK);

	// the hooks' weights: a rank holds the columns of its targets, in the order of the whole matrix
	for (int conn = 0; conn < 3; conn++) {
		int64_t e1 = 0;
		EXPECT(spice_connection_csr(one.net->context(), conn, &e1, nullptr, nullptr) == SPICE_OK);
		Int const rows = conn == 0 ? 300 : conn == 1 ? 501 : 233;
		std::vector<int64_t> off(static_cast<size_t>(rows) + 1);
		std::vector<Int32> nb(static_cast<size_t>(e1) + 1);
		std::vector<float> w1(static_cast<size_t>(e1) + 1);
		EXPECT(spice_connection_csr(one.net->context(), conn, &e1, off.data(), nb.data()) == SPICE_OK);
		EXPECT(spice_connection_synapses(one.net->context(), conn, w1.data(), e1 * 4) == SPICE_OK);
		int64_t total = 0;
		for (auto& w : r) {
			int64_t e = 0, lo = 0, hi = 0;
			EXPECT(spice_connection_csr(w.net->context(), conn, &e, nullptr, nullptr) == SPICE_OK);
			EXPECT(spice_population_range(w.net->context(), conn == 1 ? 2 : 1, &lo, &hi) == SPICE_OK);
			std::vector<float> wr(static_cast<size_t>(e) + 1);
			EXPECT(spice_connection_synapses(w.net->context(), conn, wr.data(), e * 4) == SPICE_OK);
			int64_t at = 0;
			for (int64_t i = 0; i < e1; i++)
				if (nb[static_cast<size_t>(i)] >= lo && nb[static_cast<size_t>(i)] < hi) {
					EXPECT(at < e && wr[static_cast<size_t>(at)] == w1[static_cast<size_t>(i)]);
					at++;
				}
			EXPECT(at == e);
			total += e;
		}
		EXPECT(total == e1 && e1 > 0);
	}

	long long spikes = 0;
	for (int chunk = 0; chunk < 30; chunk++) {
		one.net->run(4);
		for (auto& w : r)
			w.net->run(4);
		for (auto& w : r)
			for (Int p = 0; p < 3; p++) {
				auto got = w.net->spikes(p), want = one.net->spikes(p);
				EXPECT(got.size() == want.size());
				for (std::size_t i = 0; i < got.size(); i++)
					EXPECT(got[i] == want[i]);
			}
		spikes += static_cast<long long>(one.net->spikes(1).size() + one.net->spikes(2).size());
	}
	EXPECT(spikes > 50);
	auto a1 = state(one.A), b1 = state(one.B);
	std::vector<integrator::neuron> a2, b2;
	for (auto& w : r) {
		auto a = state(w.A), b = state(w.B);
		a2.insert(a2.end(), a.begin(), a.end());
		b2.insert(b2.end(), b.begin(), b.end());
	}
	EXPECT(a1.size() == a2.size() && b1.size() == b2.size());
	long long delivered = 0;
	for (std::size_t i = 0; i < a1.size(); i++) {
		EXPECT(a1[i].v == a2[i].v && a1[i].count == a2[i].count);
		delivered += a1[i].count;
	}
	for (std::size_t i = 0; i < b1.size(); i++)
		EXPECT(b1[i].v == b2[i].v && b1[i].count == b2[i].count);
	EXPECT(delivered > 100);
	from_to_on_two_ranks();
	std::printf("devices: rank 1 on device %d\n", device_of(1));
	std::printf("multi_rank_spec ok: %lld spikes, %lld deliveries into A\n", spikes, delivered);
	return 0;
}
