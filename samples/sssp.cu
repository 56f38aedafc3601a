// Single-source shortest paths as a spiking network (what the reference's samples/sssp.cpp demonstrates): a vertex is
// a neuron that fires in the step after its distance improved; an edge is a stateful, non-plastic synapse whose
// deliver() reads the SOURCE neuron (DeliverFromTo) and relaxes the target; edge weights are set by a per-synapse init
// hook, the source vertex by a per-population init hook.  After |V| - 1 steps every distance is final.  Prints the
// distances and checks the two the reference's sample asserts.
#include <cstdio>
#include <limits>
#include <span>

#include "spice/snn.h"

//        w3      w3
//     1 ----> 2 ----> 3
//  w1 ^                \ w1
//     0                 v
//  w1 v                 4
//     5 ------------> 6 ^ w1
//            w5
struct weighted_edge {
	Int from, to, weight;
};
static constexpr weighted_edge kGraph[] = {{0, 1, 1}, {0, 5, 1}, {1, 2, 3}, {2, 3, 3}, {3, 4, 1}, {5, 6, 5}, {6, 4, 1}};
static constexpr Int kVertices   = 7;

static Int weight_of(Int const from, Int const to) {
	for (auto const& e : kGraph)
		if (e.from == from && e.to == to)
			return e.weight;
	return 0;
}

struct vertex {
	Int source;

	struct neuron {
		Int dist   = std::numeric_limits<Int>::max();
		neuron const* via = nullptr; // the neighbour the best path arrives through
		bool improved     = false;
	};

	// per-population init hook (host): distance 0 at the source, which announces itself in the first step
	void init(std::span<neuron> all, auto& /*rng*/) {
		neuron& s  = all[static_cast<std::size_t>(source)];
		s.dist     = 0;
		s.via      = &s;
		s.improved = true;
	}

	SPICE_HD bool update(neuron& v, float /*dt*/, auto& /*rng*/) const {
		bool const announce = v.improved;
		v.improved          = false;
		return announce;
	}
};
static_assert(spice::CheckNeuron<vertex>());

struct relax {
	struct synapse {
		Int weight;
	};

	// per-synapse init hook (host)
	void init(synapse& s, Int const from, Int const to, auto& /*rng*/) const { s.weight = weight_of(from, to); }

	SPICE_HD void deliver(synapse const& s, vertex::neuron const& from, vertex::neuron& to) const {
		Int const through = from.dist + s.weight;
		if (through < to.dist) {
			to.dist     = through;
			to.via      = &from;
			to.improved = true;
		}
	}
};
static_assert(spice::CheckSynapse<relax>());

int main() {
	spice::snn net(1, 1, {1337});
	auto* vertices = net.add_population<vertex>(kVertices, {0});

	spice::adj_list edges;
	for (auto const& e : kGraph)
		edges.connect(e.from, e.to);
	net.connect<relax>(vertices, vertices, edges, 1);

	for (Int step = 0; step + 1 < vertices->size(); step++)
		net.step();

	auto const found = vertices->get_neurons();
	for (Int v = 0; v < kVertices; v++)
		std::printf("%lld%s", static_cast<long long>(found[static_cast<std::size_t>(v)].dist), v + 1 == kVertices ? "\n" : " ");
	SPICE_ASSERT(found[0].dist == 0);
	SPICE_ASSERT(found[4].dist == 7);
	return 0;
}
