// Single-source shortest paths (reference: samples/sssp.cpp): vertices are neurons that fire when
// they have found a shorter path, edges are stateful, non-plastic synapses whose deliver() reads the
// SOURCE neuron (DeliverFromTo) and whose weight comes from a per-synapse init hook; the vertices are
// set up by a per-population init hook.  Self-checking like the reference (distances of vertices 0
// and 4), and prints every vertex's distance.
#include <cstdio>
#include <limits>

#include "spice/snn.h"

using namespace spice;
using namespace spice::util;

//      3 2 3
//   1.---*---.3
//  1/         \1
// 0*           *4
//  1\         /1
//    *-------*
//    5   5   6
static Int adj_matrix[7][7] = {{0, 1, 0, 0, 0, 1, 0}, {0, 0, 3, 0, 0, 0, 0}, {0, 0, 0, 3, 0, 0, 0}, {0, 0, 0, 0, 1, 0, 0},
                               {0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 5}, {0, 0, 0, 0, 1, 0, 0}};

struct vertex {
	Int src; // the source vertex

	struct neuron {
		Int distance           = std::numeric_limits<Int>::max();
		neuron const* previous = nullptr;
		bool fire              = false;
	};

	// per-population init (runs on the host)
	void init(std::span<neuron> neurons, auto&) {
		neurons[src].distance = 0;
		neurons[src].previous = &neurons[src];
		neurons[src].fire     = true;
	}

	SPICE_HD bool update(neuron& n, float, auto&) const {
		bool const result = n.fire;
		n.fire            = false;
		return result;
	}
};
static_assert(CheckNeuron<vertex>());

struct edge {
	struct synapse {
		Int weight;
	};

	// per-synapse init (runs on the host)
	void init(synapse& syn, Int src, Int dst, auto&) const { syn.weight = adj_matrix[src][dst]; }

	SPICE_HD void deliver(synapse const& syn, vertex::neuron const& src, vertex::neuron& dst) const {
		if (src.distance + syn.weight < dst.distance) {
			dst.distance = src.distance + syn.weight;
			dst.previous = &src;
			dst.fire     = true;
		}
	}
};
static_assert(CheckSynapse<edge>());

int main() {
	snn sssp(1, 1, {1337});
	auto vertices = sssp.add_population<vertex>(7, {0});

	adj_list adj;
	for (Int src : range(7))
		for (Int dst : range(7))
			if (adj_matrix[src][dst])
				adj.connect(src, dst);

	sssp.connect<edge>(vertices, vertices, adj, 1);

	for (Int i : range(vertices->size() - 1)) {
		sssp.step();
		(void)i;
	}

	auto const result = vertices->get_neurons();
	for (Int v : range(7))
		std::printf("%lld%s", static_cast<long long>(result[v].distance), v == 6 ? "\n" : " ");
	SPICE_ASSERT(result[0].distance == 0);
	SPICE_ASSERT(result[4].distance == 7);
	return 0;
}
