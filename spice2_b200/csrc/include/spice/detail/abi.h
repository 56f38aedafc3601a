// Internal contract between the model-agnostic runtime (runtime.cu) and the kernel templates
// instantiated per user model (kernels.cuh / model_ops.cuh).  Plain structs, shared by host and
// device code.  Not part of the public C ABI (include/spice_b200.h), which only sees the two ops
// tables as opaque pointers.
#pragma once

#include <cstdint>

#include "spice/util/platform.h"

namespace spice::detail {

constexpr int kMaxWindow   = 32; // steps per launch window (also bounded by the minimum delay)
constexpr int kMaxIncoming = 8;  // connections into one population handled by one update launch
constexpr int kMaxWorld    = 16; // ranks (GPUs of one NVSwitch domain)
constexpr int kRngChunk    = 32; // consecutive neurons whose draws one thread generates per jump

struct u128 {
	std::uint64_t lo, hi;
};

// Device function applying `k` deliveries of one connection's (stateless) synapse functor to a
// neuron held in registers/local memory: k sequential calls of Syn::deliver, which is what the
// reference does event by event (synapse_population.h:118-133).
using apply_fn = void (*)(void const* functor, void* neuron, unsigned k);

// Device function applying the `n` events of one step that a stateful connection addressed to
// one neuron: `list` holds the edge indices (arrival order); it is sorted ascending — the
// reference's (source, row) order (synapse_population.h:88,99) — unless from->unordered (fast mode),
// and Syn::deliver(synapse, neuron) is called for each, reading the synapse state word-SoA
// (word w of edge e at syn[w*stride + e]).
//
// Synapses whose deliver() also takes the SOURCE neuron (concepts.h DeliverFromTo;
// synapse_population.h:125-131) find it through `from`: the source of edge e is the row that holds e
// (binary search in offsets), its state is read from a snapshot of the source population taken at the
// end of the step in which the spike was emitted (from.state == null for all other synapses).  With more than
// one rank the snapshot holds the WHOLE source population (every rank stores its slice into every peer's copy
// ahead of the step's exchange flag) and `stride` is the global population's.
struct from_ctx {
	std::uint32_t const* state; // word-SoA snapshot of the source population
	std::int64_t stride;
	std::int64_t const* offsets; // CSR row offsets of the connection
	std::int64_t n_src;
	std::int32_t unordered;      // SPICE_MODE_FAST: apply the events in arrival order (their list was filled with atomics)
	                             // instead of sorting them into the reference's (source, row) order first
};
using apply_events_fn = void (*)(void const* functor, void* neuron, std::uint32_t const* syn, std::int64_t syn_stride,
                                 std::int32_t* list, unsigned n, from_ctx const* from);

// One incoming connection as the target population's update kernel sees it.
struct incoming {
	std::uint32_t* counts;     // stateless synapses: [ring][cstride] event counters, slot = consume step % ring
	std::int64_t cstride;      // row stride of counts (local targets rounded up to 8)
	apply_fn apply;            // device function pointer (same module as the update kernel)
	void const* functor;       // device copy of the Syn object
	std::int32_t ring;
	std::int32_t zero_after_read; // 1: the producer accumulates with atomics and expects zeroed counters;
	                              // 0: the producer overwrites the whole slot (tiled delivery)
	// stateful synapses (evt_cnt != null): the events of the step that just ran, per target
	std::uint32_t* evt_cnt;        // [n_local] events addressed to each neuron (zeroed by the consumer)
	std::uint32_t const* evt_off;  // [n_local] where the neuron's events start in evt_list
	std::int32_t* evt_list;        // edge indices
	std::uint32_t const* syn;      // synapse state, word-SoA
	std::int64_t syn_stride;
	apply_events_fn apply_events;
	from_ctx from;
};

// One step of a stateful connection (launched by the runtime through spice_synapse_ops).
struct stateful_args {
	void* stream;
	void const* functor;            // device copy of the Syn object
	std::int32_t phase;             // 0 visit (catch-up + count), 1 reserve, 2 fill, 3 flush (all sources, no delivery)
	// spikes to deliver: the source population's ring slot of step time - (delay - 1)
	std::int32_t const* ring_ids;   // slot base
	std::uint32_t const* ring_cnt;  // [world] counts of the slot
	std::int64_t seg_lo[kMaxWorld];
	std::int32_t world;
	std::int64_t n_src, n_dst;      // all sources; local targets
	std::int64_t const* offsets;    // CSR (rows = all sources, local columns)
	std::int32_t const* neighbors;
	std::uint32_t* syn;             // word-SoA synapse state
	std::int64_t syn_stride;
	std::uint64_t const* dst_history; // [n_dst] 64-bit spike history of the local targets (plastic only)
	std::uint64_t* ages;              // [n_src] (last visit + 1) | delivered << 63 (synapse_population.h:89-94,137-138)
	std::int64_t time;                // snn::_time
	float dt;                         // the nominal dt (snn.cpp:19,23)
	std::uint32_t* evt_cnt;
	std::uint32_t* evt_off;
	std::uint32_t* evt_fill;
	unsigned long long* evt_cursor;
	std::int32_t* evt_list;
	std::int64_t evt_cap;
	unsigned long long* stats;        // [0] events, [1] spikes
	int* error;                       // bit 32: event list capacity exceeded
};

// Per-window random-access tables for the step streams: for every step of the window, the 128
// basis states T^i s (i < 128) of that step's xoroshiro stream folded into 4-bit lookup groups:
// nib[step][g][v] = XOR_{b in v} T^(4g+b) s, g < 32, v < 16.
struct rng_window {
	u128 const* nib; // [nsteps][32][16]
};

struct update_args {
	void* stream;           // cudaStream_t
	void const* functor;    // device copy of the Neur object
	std::uint32_t* state;   // word-SoA: word w of local neuron i at state[w * stride + i]
	std::int64_t n_local;   // neurons owned by this rank
	std::int64_t lo;        // global index of local neuron 0
	std::int64_t stride;
	std::int64_t t0;        // first step of the window (snn::_time)
	std::int32_t nsteps;
	std::int32_t cring;     // length of the incoming connections' counter rings, and the window's first slot in it
	std::int32_t cslot0;    // = t0 % cring
	std::int32_t rslot0;    // = t0 % ring (spike ring slot of the window's first step)
	float dt[kMaxWindow];   // kahan-compensated dt of each step (snn.cpp:8)
	// spike ring: ids[(step % ring) * cap + seg_base + j], cnt[(step % ring) * world + rank]
	std::int32_t* ring_ids[kMaxWorld]; // this rank's copy first ([rank]); peers' copies for direct stores
	std::uint32_t* ring_cnt;           // local counters only; published to peers after the window
	std::int64_t ring_cap;             // = global population size
	std::int32_t ring;
	std::int32_t rank, world;
	// plasticity: 64-bit spike history per local neuron (neuron_population.h:126-132) or null
	std::uint64_t* history;
	// incoming connections, in connect() order
	std::int32_t n_in;
	incoming in[kMaxIncoming];
	// random access into the step streams
	rng_window rng;
	u128 const* jump_poly;  // per chunk of kRngChunk local neurons: x^(offset of chunk start) mod charpoly
	int* error;             // bit 64: a neuron's update() drew more values than its rng_draws declares
};

struct export_args {
	void* stream;
	std::uint32_t const* state;
	std::int64_t n_local, stride;
	std::int64_t t_next; // the step whose pending deliveries must be folded in (= snn::_time)
	std::int32_t n_in;
	incoming in[kMaxIncoming];
	void* out_aos; // device buffer, n_local * sizeof(neuron)
};

struct import_args {
	void* stream;
	std::uint32_t* state;
	std::int64_t n_local, stride;
	void const* in_aos;
	// what export_args folds into the copy it hands out is pending no more once that copy comes back: the events of
	// step t_next are cleared, so a get / modify / set round trip applies them once (neuron_population.h:142-145 hands
	// out the live state, where they have been applied already)
	std::int64_t t_next;
	std::int32_t n_in;
	incoming in[kMaxIncoming];
};
}

// ---- the two ops tables behind the opaque pointers of include/spice_b200.h ---------------------
extern "C" {
struct spice_neuron_ops {
	std::uint32_t abi_version;
	char const* name;
	std::uint32_t neuron_bytes;  // sizeof(Neur::neuron); 0 for stateless neurons
	std::uint32_t functor_bytes; // sizeof(Neur)
	std::uint32_t rng_draws;     // engine draws one update() consumes (compile-time constant)
	std::uint32_t burns_seed;    // stateful adapters consume one seed++ (neuron_population.h:60)
	// host: fill `out` (n * neuron_bytes) with default-constructed neurons and run the model's
	// init hook, if any, with an engine seeded by (seed_lo, seed_hi) (neuron_population.h:60-67)
	void (*init_host)(void const* functor, void* out, std::int64_t n, std::uint64_t seed_lo, std::uint64_t seed_hi);
	int (*launch_update)(spice::detail::update_args const*);
	int (*launch_export)(spice::detail::export_args const*);
	int (*launch_import)(spice::detail::import_args const*);
};

struct spice_synapse_ops {
	std::uint32_t abi_version;
	char const* name;
	std::uint32_t synapse_bytes; // sizeof(Syn::synapse); 0 for stateless synapses
	std::uint32_t functor_bytes; // sizeof(Syn)
	std::uint32_t dst_neuron_bytes;
	std::uint32_t plastic;         // has update()/skip()
	std::uint32_t deliver_from_to; // deliver takes the source neuron (see from_ctx; more than one rank: snapshots in the exchange region)
	// device function pointer of apply<Syn, DstNeur>, fetched from the module that holds the kernels
	int (*get_apply)(spice::detail::apply_fn* out);
	// stateful synapses: default-construct `n_edges` synapses (AoS) and run the model's init hook, if any,
	// over the CSR in row order with an engine seeded by (seed_lo, seed_hi) (synapse_population.h:34-41)
	std::uint32_t per_synapse_init; // the hook exists (it consumes one seed++)
	void (*init_host)(void const* functor, void* out_aos, std::int64_t const* offsets, std::int32_t const* neighbors,
	                  std::int64_t n_src, std::uint64_t seed_lo, std::uint64_t seed_hi);
	int (*get_apply_events)(spice::detail::apply_events_fn* out);
	int (*launch_stateful)(spice::detail::stateful_args const*);
	// stateless synapses: the target model's update kernel with this synapse's deliver() inlined; the
	// same address for every source type, so the runtime can tell that all incoming connections of a
	// population agree (null: not available)
	int (*launch_update_fused)(spice::detail::update_args const*);
};
}
