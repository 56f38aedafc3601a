// Connection topologies, as descriptors (reference: spice/include/spice/topology.h:24-58).
//
// In the reference a Topology generates its adjacency on the host (topology.cpp).  Here it only
// describes the adjacency; snn::connect hands the description to the backend, which generates
// fixed_probability adjacencies on the GPU, bit-exact with fixed_probability::generate
// (topology.cpp:80-112), and sorts adj_list edges into CSR as adj_list::generate does
// (topology.cpp:63-71).
#pragma once

#include <limits>
#include <vector>

#include "spice/util/assert.h"
#include "spice/util/platform.h"

namespace spice {
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

	std::vector<Int32> const& sources() const { return _src; }
	std::vector<Int32> const& targets() const { return _dst; }

private:
	std::vector<Int32> _src, _dst;
};

class fixed_probability : public Topology {
public:
	explicit fixed_probability(double const p) : _p(p) { SPICE_PRE(0 <= p && p <= 1); }

	Int size() const override; // src_count * max_degree (topology.cpp:75-78); defined in snn.h
	double p() const { return _p; }

private:
	double _p;
};
}
