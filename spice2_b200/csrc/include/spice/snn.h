// class spice::snn — the reference's C++20 API (spice/include/spice/snn.h:16-75), kept intact as
// the drop-in boundary, over the B200 backend.
//
//     snn net(dt, max_delay, {1337});
//     auto P = net.add_population<poisson>(N / 2);
//     auto E = net.add_population<lif>(N * 4 / 10);
//     net.connect<fixed_weight>(P, E, fixed_probability(0.1), delay, {2.0 / N});
//     net.step();
//     E->spikes(0);
//
// Differences a user sees:
//   * the translation unit is compiled by nvcc (-std=c++20 -fmad=false, sm_100a): add_population
//     and connect instantiate the simulation kernels for the user's functors right here;
//   * functor members that run in kernels carry SPICE_HD; neurons that draw declare rng_draws;
//   * step() enqueues work; results are observable through spikes()/get_neurons(), which
//     synchronise.  run(n) advances n steps with one launch window per min-delay steps;
//   * snn::spikes(i) (named by the north star) = spikes(0) of the i-th population added.
// Errors keep the reference's convention: std::logic_error("Assertion failed (file:line): cond").
#pragma once

#include <algorithm>
#include <cmath>
#include <memory>
#include <span>
#include <stdexcept>
#include <utility>
#include <vector>

#include "spice/concepts.h"
#include "spice/detail/model_ops.cuh"
#include "spice/topology.h"
#include "spice/util/range.h"
#include "spice/util/numeric.h"
#include "spice/util/random.h"
#include "spice_b200.h"

namespace spice {
namespace detail {
inline void check(spice_ctx* ctx, int rc) {
	if (rc != SPICE_OK)
		throw std::logic_error(spice_last_error(ctx));
}

// type-erased view (reference: detail::NeuronPopulation, neuron_population.h:17-25)
struct NeuronPopulation {
	virtual ~NeuronPopulation() = default;
	virtual Int size() const                             = 0;
	virtual std::span<Int32 const> spikes(Int age) const = 0;
	virtual int index() const                            = 0;
};

template <class T>
struct state_cache {
	std::vector<T> values;
};
template <>
struct state_cache<void> {};

template <Neuron Neur>
class neuron_population : public NeuronPopulation {
public:
	neuron_population(spice_ctx* ctx, Neur neuron, Int const size) : _ctx(ctx), _size(size), _host(std::move(neuron)) {
		if constexpr (PerPopulationUpdate<Neur>) // spikes come from the host functor, step by step (neuron_population.h:86-101)
			check(ctx, spice_add_host_population(ctx, size, &neuron_population::host_update, this, &_index));
		else
			check(ctx, spice_add_population(ctx, neuron_ops<Neur>(), size, &_host, &_index));
	}
	neuron_population(neuron_population const&)            = delete; // the runtime holds `this`
	neuron_population& operator=(neuron_population const&) = delete;

	Int size() const override { return _size; }
	int index() const override { return _index; }

	// spikes emitted `age` steps ago, ascending (neuron_population.h:147-153); valid until the next step()
	std::span<Int32 const> spikes(Int age) const override {
		Int32 const* ids = nullptr;
		int64_t n        = 0;
		check(_ctx, spice_spikes(_ctx, _index, age, &ids, &n));
		return {ids, static_cast<std::size_t>(n)};
	}

	// copy of this rank's neurons (neuron_population.h:142-145), refreshed on every call.  The reference hands out a span
	// over the live state; the state lives on the device here, so the span is read-only (a write into a copy would be lost
	// silently: it does not compile instead) and set_neurons() carries changes back.
	auto get_neurons() {
		static_assert(StatefulNeuron<Neur>, "Can only return collections of stateful neurons.");
		using N = neuron_traits_t<Neur>;
		int64_t lo = 0, hi = 0;
		check(_ctx, spice_population_range(_ctx, _index, &lo, &hi));
		_cache.values.resize(static_cast<std::size_t>(hi - lo));
		check(_ctx, spice_neurons(_ctx, _index, _cache.values.data(), static_cast<int64_t>(_cache.values.size() * sizeof(N))));
		return std::span<N const>(_cache.values);
	}
	// overwrite this rank's neurons (as many as get_neurons() returns); deliveries still pending for the next step are
	// part of what get_neurons() returned and are not applied a second time
	void set_neurons(std::span<neuron_traits_t<Neur> const> values) {
		static_assert(StatefulNeuron<Neur>, "Can only set collections of stateful neurons.");
		check(_ctx, spice_set_neurons(_ctx, _index, values.data(), static_cast<int64_t>(values.size_bytes())));
	}

private:
	// engine over the step's stream that counts its draws
	struct counting_engine {
		using result_type = UInt;
		util::xoroshiro64_128p g;
		int64_t* draws;
		static constexpr result_type min() { return util::xoroshiro64_128p::min(); }
		static constexpr result_type max() { return util::xoroshiro64_128p::max(); }
		result_type operator()() {
			++*draws;
			return g();
		}
	};
	static int64_t host_update(void* user, float dt, uint64_t seed_lo, uint64_t seed_hi, uint64_t rng_offset, int32_t* ids_out,
	                           int64_t capacity, int64_t* draws_out) {
		if constexpr (PerPopulationUpdate<Neur>) {
			auto* self = static_cast<neuron_population*>(user);
			counting_engine rng{util::xoroshiro64_128p(seed_lo, seed_hi), draws_out};
			for (uint64_t i = 0; i < rng_offset; i++) // populations before this one drew first (snn.cpp:12-15)
				rng.g();
			self->_out.clear();
			self->_host.update(dt, rng, self->_out);
			if (static_cast<int64_t>(self->_out.size()) > capacity)
				return -1;
			std::copy(self->_out.begin(), self->_out.end(), ids_out);
			return static_cast<int64_t>(self->_out.size());
		} else {
			(void)user, (void)dt, (void)seed_lo, (void)seed_hi, (void)rng_offset, (void)ids_out, (void)capacity, (void)draws_out;
			return -1;
		}
	}

	spice_ctx* _ctx;
	Int _size;
	int _index = -1;
	Neur _host;               // the functor; host-fed populations keep calling it
	std::vector<Int32> _out;  // its output buffer
	[[no_unique_address]] state_cache<neuron_traits_t<Neur>> _cache;
};
}

inline Int fixed_probability::size() const { return src_count * spice_fixed_probability_max_degree(dst_count, _p); }

inline void fixed_probability::generate(std::span<Int> offsets, std::span<Int32> neighbors, util::seed_seq const& seed) {
	SPICE_PRE(static_cast<Int>(offsets.size()) > src_count);
	SPICE_PRE(static_cast<Int>(neighbors.size()) >= size());
	spice_adjacency* a = nullptr;
	if (spice_fixed_probability_generate(device, src_count, dst_count, _p, seed.seed().lo, seed.seed().hi, 0, dst_count, &a) != SPICE_OK)
		throw std::logic_error(spice_last_error(nullptr));
	int const rc = spice_adjacency_copy(a, offsets.data(), neighbors.data());
	spice_adjacency_destroy(a);
	if (rc != SPICE_OK)
		throw std::logic_error(spice_last_error(nullptr));
}

class snn {
public:
	snn(float const dt, float const max_delay, util::seed_seq seed, int device = 0, int rank = 0, int world = 1,
	    int mode = SPICE_MODE_DETERMINISTIC) :
	_dt(dt) {
		spice_ctx* ctx = nullptr;
		int const rc   = spice_ctx_create_seeded(&ctx, device, dt, max_delay, seed.seed().lo, seed.seed().hi, rank, world, mode);
		if (rc != SPICE_OK)
			throw std::logic_error(spice_last_error(nullptr));
		_ctx.reset(ctx);
	}

	template <Neuron Neur>
	detail::neuron_population<Neur>* add_population(Int const size, Neur neur = {}) {
		_neurons.push_back(std::make_unique<detail::neuron_population<Neur>>(_ctx.get(), std::move(neur), size));
		return static_cast<detail::neuron_population<Neur>*>(_neurons.back().get());
	}

	// Not in the reference (one address space): the population's target ranges over the ranks, bounds[world + 1], instead of
	// equal widths — e.g. the in-degree-balanced ranges of spice_balance_ranges for a non-uniform topology.
	template <Neuron Neur>
	detail::neuron_population<Neur>* add_population(Int const size, Neur neur, std::span<Int const> bounds) {
		static_assert(sizeof(Int) == sizeof(int64_t));
		detail::check(_ctx.get(), spice_set_next_partition(_ctx.get(), reinterpret_cast<int64_t const*>(bounds.data())));
		return add_population<Neur>(size, std::move(neur));
	}

	template <class Syn, Neuron SrcNeur, StatefulNeuron DstNeur>
	requires Synapse<Syn, SrcNeur, DstNeur>
	void connect(detail::neuron_population<SrcNeur>* source, detail::neuron_population<DstNeur>* target, Topology& c,
	             float const delay, Syn syn = {}) {
		c(source->size(), target->size());
		auto const* ops = detail::synapse_ops<Syn, SrcNeur, DstNeur>();
		if (auto* fp = dynamic_cast<fixed_probability*>(&c))
			detail::check(_ctx.get(), spice_connect_fixed_probability(_ctx.get(), ops, source->index(), target->index(), fp->p(),
			                                                          delay, &syn, nullptr));
		else if (auto* adj = dynamic_cast<adj_list*>(&c))
			detail::check(_ctx.get(), spice_connect_adj_list(_ctx.get(), ops, source->index(), target->index(), adj->sources().data(),
			                                                 adj->targets().data(), adj->size(), delay, &syn, nullptr));
		else {
			// a user-defined Topology: run its generate() on the host with the seed the connection is about to consume
			// (synapse_population.h:31 `_graph(c, seed++)`, csr.h:69-77), then upload the rows like an adj_list's.  Rows
			// arrive sorted by target; that is the order generate() emitted whenever it emitted ascending targets.
			uint64_t sd[2];
			detail::check(_ctx.get(), spice_ctx_seed(_ctx.get(), sd));
			std::vector<Int> offsets(static_cast<std::size_t>(c.src_count > 0 ? c.src_count + 1 : 0));
			std::vector<Int32> neighbors(static_cast<std::size_t>(c.size()));
			c.generate(offsets, neighbors, util::seed_seq(UInt128{sd[0], sd[1]}));
			SPICE_INV(std::is_sorted(offsets.begin(), offsets.end()));
			Int const edges = offsets.empty() ? 0 : offsets.back();
			SPICE_INV(edges <= static_cast<Int>(neighbors.size()));
			std::vector<Int32> srcs(static_cast<std::size_t>(edges));
			for (Int row = 0; row < c.src_count; row++)
				std::fill(srcs.begin() + offsets[static_cast<std::size_t>(row)], srcs.begin() + offsets[static_cast<std::size_t>(row) + 1],
				          static_cast<Int32>(row));
			detail::check(_ctx.get(), spice_connect_adj_list(_ctx.get(), ops, source->index(), target->index(), srcs.data(),
			                                                 neighbors.data(), edges, delay, &syn, nullptr));
		}
	}

	template <class Syn, Neuron SrcNeur, StatefulNeuron DstNeur>
	requires Synapse<Syn, SrcNeur, DstNeur>
	void connect(detail::neuron_population<SrcNeur>* source, detail::neuron_population<DstNeur>* target, Topology&& c,
	             float const delay, Syn syn = {}) {
		connect<Syn, SrcNeur, DstNeur>(source, target, c, delay, std::move(syn));
	}

	// snn::step() (spice/src/snn.cpp:7-28)
	void step() { run(1); }
	// n steps, one launch window per min-delay steps
	void run(Int n) { detail::check(_ctx.get(), spice_run(_ctx.get(), n)); }
	void sync() { detail::check(_ctx.get(), spice_sync(_ctx.get())); }

	// spikes of the i-th population added, emitted in the step that just ran
	std::span<Int32 const> spikes(Int i) const { return _neurons.at(static_cast<std::size_t>(i))->spikes(0); }

	spice_ctx* context() { return _ctx.get(); }

private:
	struct ctx_deleter {
		void operator()(spice_ctx* c) const { spice_ctx_destroy(c); }
	};
	float _dt;
	std::unique_ptr<spice_ctx, ctx_deleter> _ctx;
	std::vector<std::unique_ptr<detail::NeuronPopulation>> _neurons;
};
}
