// Connection topologies (reference: spice/include/spice/topology.h:11-58, spice/src/topology.cpp).
//
// In the reference a Topology generates its adjacency on the host (topology.cpp).  Here the two
// built-in topologies are descriptors: snn::connect hands the description to the backend, which
// generates fixed_probability adjacencies on the GPU, bit-exact with fixed_probability::generate
// (topology.cpp:80-112), and sorts adj_list edges into CSR as adj_list::generate does
// (topology.cpp:63-71).  The reference's extension point is kept: a user-defined Topology
// overrides generate(edge_stream&, seed) (or the span form), snn::connect runs it on the host with
// the connection's seed and uploads the rows.  generate() of the built-in topologies is callable
// directly as well, as bench/connectivity.cpp does (fixed_probability: on the GPU, copied back).
#pragma once

#include <algorithm>
#include <limits>
#include <span>
#include <utility>
#include <vector>

#include "spice/util/assert.h"
#include "spice/util/platform.h"
#include "spice/util/random.h"

namespace spice {
// Writes edges, ascending in src, into CSR arrays (topology.h:11-22, topology.cpp:12-32).
class edge_stream {
public:
	edge_stream(std::span<Int> offsets, std::span<Int32> neighbors) : _offsets(offsets), _neighbors(neighbors) {}

	edge_stream& operator<<(std::pair<Int32, Int32> const edge) {
		SPICE_PRE(static_cast<UInt>(_src) < _offsets.size());
		SPICE_PRE(static_cast<UInt>(_dst) < _neighbors.size());
		SPICE_PRE(static_cast<UInt>(edge.first) < _offsets.size());
		while (_src <= edge.first) // rows up to and including the edge's start here
			_offsets[static_cast<std::size_t>(_src++)] = _dst;
		_neighbors[static_cast<std::size_t>(_dst++)] = edge.second;
		return *this;
	}

	// closes the open row; the stream starts over
	void flush() {
		SPICE_INV(static_cast<UInt>(_src) < _offsets.size());
		_offsets[static_cast<std::size_t>(_src)] = _dst;
		_src = _dst = 0;
	}

private:
	std::span<Int> _offsets;
	std::span<Int32> _neighbors;
	Int _src = 0;
	Int _dst = 0;
};

struct Topology {
	Int src_count = 0;
	Int dst_count = 0;

	virtual ~Topology() = default;

	// bind the topology to the population sizes it connects (topology.cpp:34-41)
	Topology& operator()(Int const src_count_, Int const dst_count_) {
		SPICE_INV(0 <= src_count_ && src_count_ < std::numeric_limits<Int32>::max());
		SPICE_INV(0 <= dst_count_ && dst_count_ < std::numeric_limits<Int32>::max());
		src_count = src_count_;
		dst_count = dst_count_;
		return *this;
	}

	// upper bound on the number of edges
	virtual Int size() const = 0;

	// topology.cpp:43-55: a subclass implements one of the two
	virtual void generate(edge_stream&, util::seed_seq const&) {
		SPICE_PRE(false && "Topology subclasses must implement generate(edge_stream, seed_seq)");
	}
	virtual void generate(std::span<Int> offsets, std::span<Int32> neighbors, util::seed_seq const& seed) {
		SPICE_PRE(static_cast<Int>(offsets.size()) > src_count);
		SPICE_PRE(static_cast<Int>(neighbors.size()) >= size());
		edge_stream es(offsets, neighbors);
		generate(es, seed);
		es.flush();
	}
};

class adj_list : public Topology {
public:
	void connect(Int const src, Int const dst) {
		SPICE_PRE(0 <= src && src < std::numeric_limits<Int32>::max());
		SPICE_PRE(0 <= dst && dst < std::numeric_limits<Int32>::max());
		_src.push_back(static_cast<Int32>(src));
		_dst.push_back(static_cast<Int32>(dst));
	}

	Int size() const override { return static_cast<Int>(_src.size()); }

	// topology.cpp:63-71: edges sorted by (src, dst)
	using Topology::generate;
	void generate(edge_stream& stream, util::seed_seq const&) override {
		std::vector<UInt> packed(_src.size());
		for (std::size_t i = 0; i < _src.size(); i++)
			packed[i] = static_cast<UInt>(_src[i]) << 32 | static_cast<UInt32>(_dst[i]);
		std::sort(packed.begin(), packed.end());
		for (UInt const c : packed)
			stream << std::pair{static_cast<Int32>(c >> 32), static_cast<Int32>(c & 0xffffffffu)};
	}

	std::vector<Int32> const& sources() const { return _src; }
	std::vector<Int32> const& targets() const { return _dst; }

private:
	std::vector<Int32> _src, _dst;
};

class fixed_probability : public Topology {
public:
	explicit fixed_probability(double const p) : _p(p) { SPICE_PRE(0 <= p && p <= 1); }

	Int size() const override; // src_count * max_degree (topology.cpp:75-78); defined in snn.h
	// topology.cpp:80-112, generated on GPU `device` and copied back; defined in snn.h
	using Topology::generate;
	void generate(std::span<Int> offsets, std::span<Int32> neighbors, util::seed_seq const& seed) override;
	int device = 0;
	double p() const { return _p; }

private:
	double _p;
};
}
