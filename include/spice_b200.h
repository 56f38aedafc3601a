/* spice_b200.h — C ABI of the B200 backend for Spice's per-timestep simulation loop and
 * synapse generation.
 *
 * The reference (denniskb/spice2 @ f5e57eb) has no FFI: its boundary is the C++20 template API
 * of `class spice::snn` plus the functor concepts (spice/include/spice/snn.h:16-75,
 * spice/include/spice/concepts.h:11-104).  This backend keeps that C++ API intact
 * (spice2_b200/csrc/include/spice/snn.h) and puts this thin C ABI underneath it: plain
 * pointers, sizes and opaque handles, `int` status returns, no exceptions and no torch / C++
 * types across the boundary.  The C++ facade turns a non-zero status into the reference's
 * error behaviour, `throw std::logic_error("Assertion failed (file:line): cond")`
 * (spice/include/spice/util/assert.h:3-17, spice/src/util/assert.cpp:8-15).
 *
 * Each entry point cites the reference interface it stands in for.  INTEGRATION.md shows the
 * binding a maintainer of the reference would add on their side.
 *
 * There is no CPU fallback: every call that computes needs a CUDA device (sm_100a).
 */
#ifndef SPICE_B200_H
#define SPICE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPICE_B200_ABI_VERSION 1

#if defined(__GNUC__)
#define SPICE_API __attribute__((visibility("default")))
#else
#define SPICE_API
#endif

/* status codes */
enum {
	SPICE_OK                = 0,
	SPICE_ERR_PRECONDITION  = 1, /* a SPICE_PRE of the reference would have fired */
	SPICE_ERR_CUDA          = 2,
	SPICE_ERR_UNSUPPORTED   = 3,
	SPICE_ERR_INTERNAL      = 4, /* e.g. the generator's self-check failed */
	SPICE_ERR_NO_DEVICE     = 5
};

/* delivery modes (new in this backend) */
enum {
	/* per-(slot, connection, target) integer event counts, applied by the target's owner thread
	 * in connect() order: bit-exact with the reference's sequential float accumulation for
	 * stateless synapses.  Stateful synapses are applied in (connection, source rank, row) order. */
	SPICE_MODE_DETERMINISTIC = 0,
	/* stateful synapses: the per-target event lists are filled with atomics and applied in arrival order, without the
	 * sort into the reference's (source, row) order, so the order of a target's float accumulation is not defined:
	 * membrane potentials agree with the deterministic mode to float rounding of the sums, firing rates within the
	 * tolerance tests/test_gpu_sim.py::test_fast_mode_tolerances states.  Stateless synapses use the same integer
	 * counters as the deterministic mode (already order-free and exact). */
	SPICE_MODE_FAST = 1
};

typedef struct spice_ctx spice_ctx;

/* Type-erased model tables.  They are produced by template instantiations in a .cu translation
 * unit that sees the user's functor types (spice/detail/model_ops.cuh); the facade obtains them
 * implicitly from add_population<N>() / connect<S>(), and spice_builtin_*() returns the tables
 * of the sample models compiled into the library. */
typedef struct spice_neuron_ops spice_neuron_ops;
typedef struct spice_synapse_ops spice_synapse_ops;

/* ---- context -------------------------------------------------------------------------------
 * snn::snn(float dt, float max_delay, util::seed_seq seed)          (snn.h:18-19)
 * `seed_words` is the initializer list handed to util::seed_seq      (random.h:149-152).
 * rank/world: target-partitioned multi-GPU execution (one process per GPU); this context owns
 * neurons [size*rank/world, size*(rank+1)/world) of every population and all their incoming
 * synapses.  world == 1 for single-GPU. */
SPICE_API int spice_ctx_create(spice_ctx** out, int device, float dt, float max_delay, uint32_t const* seed_words,
                     int n_seed_words, int rank, int world, int mode);
/* same, from an already-hashed 128-bit seed (util::seed_seq::seed(), random.h:161) — what the C++
 * facade holds when the user passes `{1337}` through the reference's constructor signature */
SPICE_API int spice_ctx_create_seeded(spice_ctx** out, int device, float dt, float max_delay, uint64_t seed_lo,
                                      uint64_t seed_hi, int rank, int world, int mode);
SPICE_API int spice_ctx_destroy(spice_ctx* ctx);
/* message of the last failure on this context (or of the last failed create when ctx == NULL) */
SPICE_API char const* spice_last_error(spice_ctx const* ctx);
/* run all work on this CUDA stream (a cudaStream_t), e.g. torch's current stream.  Default: a
 * stream the context creates. */
SPICE_API int spice_ctx_set_stream(spice_ctx* ctx, void* cuda_stream);

/* ---- populations ---------------------------------------------------------------------------
 * snn::add_population<Neur>(Int size, Neur neur = {})               (snn.h:21-27)
 * `functor` points at a Neur object (ops->functor_bytes bytes), copied into the population. */
SPICE_API int spice_add_population(spice_ctx* ctx, spice_neuron_ops const* ops, int64_t size, void const* functor,
                         int* pop_out);
/* snn::add_population for a neuron with a per-population update()          (concepts.h:46-57,
 * neuron_population.h:86-101; samples/external_input.cpp): the population's spikes come from the
 * host.  When a step is enqueued the runtime calls update(user, dt, seed_lo, seed_hi, rng_offset,
 * ids_out, capacity, &draws): it writes the population-relative ids (0 <= id < size) of the neurons
 * that fire in this step and returns how many (<= capacity = size; < 0: failure).  (seed, rng_offset)
 * address the step's random stream — the engine of snn.cpp:12 advanced by the draws of the
 * populations added before this one (and by the host functors that ran before this one in the same step);
 * *draws_out = how many values this call took.  A functor that draws must come after every device
 * population that draws (those jump into the stream at positions fixed when the network is built).
 * With more than one rank the population is replicated: every rank calls its own copy of the functor,
 * which must emit the same spikes everywhere. */
typedef int64_t (*spice_host_update_fn)(void* user, float dt, uint64_t seed_lo, uint64_t seed_hi, uint64_t rng_offset,
                                        int32_t* ids_out, int64_t capacity, int64_t* draws_out);
SPICE_API int spice_add_host_population(spice_ctx* ctx, int64_t size, spice_host_update_fn update, void* user, int* pop_out);
/* Not in the reference (it has one address space): the target ranges of the NEXT population added with
 * spice_add_population, bounds[world + 1], bounds[0] = 0, bounds[world] = size, non-decreasing; rank r owns neurons
 * [bounds[r], bounds[r + 1]) and all their incoming synapses.  Every rank must pass the same bounds.  NULL (the default)
 * = equal widths, which balances fixed_probability.  For non-uniform topologies (adj_list, user-defined Topology) pass the
 * ranges spice_balance_ranges computes from the targets' in-degrees: static synapse-count load balancing. */
SPICE_API int spice_set_next_partition(spice_ctx* ctx, int64_t const* bounds);
/* bounds[world + 1] such that every range holds about the same sum of weight[i] + 1 (weight = in-degree of target i summed
 * over the connections into the population; + 1 = the neuron's own update).  Host arithmetic, no device needed. */
SPICE_API int spice_balance_ranges(int64_t const* weight, int64_t n, int world, int64_t* bounds);
/* NeuronPopulation::size()                                           (neuron_population.h:114) */
SPICE_API int64_t spice_population_size(spice_ctx const* ctx, int pop);
/* the contiguous range of the population this rank owns */
SPICE_API int spice_population_range(spice_ctx const* ctx, int pop, int64_t* lo, int64_t* hi);

/* ---- connections ---------------------------------------------------------------------------
 * snn::connect<Syn>(source, target, fixed_probability(p), delay, syn) (snn.h:29-56)
 * PRE 0 <= p <= 1 (topology.cpp:73); PRE 1 <= round(delay/dt) <= max_delay (snn.h:35-38). */
SPICE_API int spice_connect_fixed_probability(spice_ctx* ctx, spice_synapse_ops const* ops, int src_pop, int dst_pop,
                                    double p, float delay, void const* functor, int* conn_out);
/* The same connection drawn by this backend's counter-based generator (spice_fixed_probability_generate_fast below):
 * independent Bernoulli(p) per (source, target) pair, rows generated in parallel at write bandwidth.  NOT the reference's
 * matrix for the same seed (its stream is sequential): for networks whose results are compared statistically.  The
 * connection consumes the same seed increment, so everything else of the network keeps its streams. */
SPICE_API int spice_connect_fixed_probability_fast(spice_ctx* ctx, spice_synapse_ops const* ops, int src_pop, int dst_pop,
                                         double p, float delay, void const* functor, int* conn_out);
/* snn::connect<Syn>(source, target, adj_list, delay, syn)            (snn.h:29-56, topology.h:37-46)
 * edges: n_edges (src, dst) pairs; sorted by (src,dst) into CSR as adj_list::generate does
 * (topology.cpp:63-71). */
SPICE_API int spice_connect_adj_list(spice_ctx* ctx, spice_synapse_ops const* ops, int src_pop, int dst_pop,
                           int32_t const* edges_src, int32_t const* edges_dst, int64_t n_edges, float delay,
                           void const* functor, int* conn_out);
/* csr<T>: _offsets / _neighbors (csr.h:97-99).  Copies the connection's CSR to host memory:
 * offsets has src_size+1 entries; neighbors holds this rank's column slice with LOCAL column
 * indices (dst - lo).  Pass NULL to query sizes only. */
SPICE_API int spice_connection_csr(spice_ctx* ctx, int conn, int64_t* n_edges_out, int64_t* offsets_out,
                         int32_t* neighbors_out);
/* csr<T>::_edges (csr.h:99): per-synapse state of a stateful connection, AoS, parallel to
 * neighbors. */
SPICE_API int spice_connection_synapses(spice_ctx* ctx, int conn, void* out, int64_t bytes);

/* ---- stepping ------------------------------------------------------------------------------
 * snn::step()                                                        (snn.cpp:7-28), n times.
 * Asynchronous with respect to the host; results are observable through the readout calls
 * below, which synchronise.  Steps are executed in windows of at most the minimum connection
 * delay (no spike emitted inside a window can be consumed inside it). */
SPICE_API int spice_run(spice_ctx* ctx, int64_t n_steps);
SPICE_API int spice_sync(spice_ctx* ctx);
SPICE_API int64_t spice_time(spice_ctx const* ctx); /* snn::_time */

/* NeuronPopulation::spikes(age)                                      (neuron_population.h:147-153)
 * ids_out receives a pointer into context-owned pinned host memory holding the (global,
 * population-relative) ids of the neurons that fired `age` steps ago, ascending; valid until
 * the next spice_run().  PRE 0 <= age < min(steps run, max_delay). */
SPICE_API int spice_spikes(spice_ctx* ctx, int pop, int64_t age, int32_t const** ids_out, int64_t* n_out);
/* neuron_population::get_neurons()                                   (neuron_population.h:142-145)
 * copies this rank's neurons as an array of Neur::neuron (AoS, as the reference lays them out). */
SPICE_API int spice_neurons(spice_ctx* ctx, int pop, void* out, int64_t bytes);
/* overwrite this rank's neuron state (the reference exposes a mutable span) */
SPICE_API int spice_set_neurons(spice_ctx* ctx, int pop, void const* in, int64_t bytes);

/* Spike sink: batched readout of every step's spike lists (the per-step sink of the samples,
 * samples/matplot.cpp:96-134, without a device sync per step).  While enabled, each window's
 * lists are sorted on the device and written into a ring in page-locked host memory; a readout
 * waits only for the steps it takes, so the host can drain batch k while the device runs batch
 * k + 1.  spice_raster_size waits for the first max_steps unread steps (<= 0: all steps issued)
 * and reports how many steps / ids they hold; spice_raster_read copies them out and frees their
 * ring space: counts[step * n_pops + pop] and the concatenated ascending ids in (step, pop)
 * order.  An unread log that outgrows the ring is an error reported by the next readout. */
SPICE_API int spice_raster_enable(spice_ctx* ctx, int enable);
SPICE_API int spice_raster_size(spice_ctx* ctx, int64_t max_steps, int64_t* n_steps_out, int64_t* n_ids_out);
SPICE_API int spice_raster_read(spice_ctx* ctx, int64_t n_steps, int64_t* counts_out, int32_t* ids_out);

/* counters: synaptic events (Syn::deliver invocations, synapse_population.h:118-133) and spikes
 * processed by this rank since creation */
SPICE_API int spice_stats(spice_ctx* ctx, int64_t* synaptic_events, int64_t* spikes_delivered, int64_t* kernel_launches);

/* Phase timing with CUDA events recorded on the context's stream around each window's update
 * and delivery launches (measurement only; bench.py's roofline numbers come from here).
 * enable = n > 1: around every n-th window only (four timed events per window cost ~3 % of a 270 us
 * window; sampled, the timed region carries a quarter of that).  spice_profile_read synchronises,
 * returns the sums over the marked windows since the last read (and how many there were) and clears them. */
SPICE_API int spice_profile_enable(spice_ctx* ctx, int enable);
/* launch windows enqueued so far (one per min-delay steps; one per step for networks with stateful synapses) */
SPICE_API int64_t spice_windows_run(spice_ctx const* ctx);
SPICE_API int spice_profile_read(spice_ctx* ctx, double* update_ms, double* deliver_ms, double* exchange_ms,
                                 int64_t* windows);

/* ---- multi-GPU spike exchange (new) ---------------------------------------------
 * One process per GPU.  Every rank keeps a copy of each population's spike ring in one device
 * allocation (the "exchange region"); the update kernels store the spikes they emit straight
 * into every peer's copy over NVLink (peer stores), and a per-window flag tells the peers when a
 * window's spikes are complete — an all-gather of spike ids fused into the producing kernel,
 * once per min-delay window.  Setup: every rank calls spice_ctx_finalize() after its last
 * connect(), exports its handle, the caller all-gathers the handles (torch.distributed, MPI, a
 * file ...) and hands the world's handles, rank-major, to spice_ctx_set_peers().  Handles of
 * contexts living in the same process (tests) are resolved without CUDA IPC. */
SPICE_API int spice_ctx_finalize(spice_ctx* ctx);
SPICE_API int spice_ctx_peer_handle(spice_ctx* ctx, void* out, int64_t* bytes); /* out == NULL: size query */
SPICE_API int spice_ctx_set_peers(spice_ctx* ctx, void const* handles, int64_t bytes_each);

/* ---- standalone synapse generation ----------------------------------------------------------
 * fixed_probability::size()                                          (topology.cpp:75-78) */
SPICE_API int64_t spice_fixed_probability_max_degree(int64_t dst_count, double p);
/* fixed_probability::generate(offsets, neighbors, seed)              (topology.cpp:80-112)
 * on the GPU, bit-exact with the reference for the same seed.  seed_lo/hi = seed_seq::seed().
 * Keeps columns [col_lo, col_hi) (pass 0, dst_count for all) with local indices.  Results stay
 * on the device (returned as device pointers owned by the handle) unless host buffers are
 * given.  offsets_host: src_count+1 entries; neighbors_host: at least *n_edges_out entries
 * (query first with NULL buffers, or size by spice_fixed_probability_max_degree). */
typedef struct spice_adjacency spice_adjacency;
SPICE_API int spice_fixed_probability_generate(int device, int64_t src_count, int64_t dst_count, double p, uint64_t seed_lo,
                                     uint64_t seed_hi, int64_t col_lo, int64_t col_hi, spice_adjacency** out);
/* Not in the reference: the generator its sampler becomes when every row may use its own engine (seed_seq::stream(id),
 * random.h:169, which the reference defines and never uses): engine (row * 32 + lane), geometric skips between connected
 * targets, 32 per warp iteration, coalesced stores.  Same accessors; rows ascending, no duplicates, degree ~ Binomial(dst, p). */
SPICE_API int spice_fixed_probability_generate_fast(int device, int64_t src_count, int64_t dst_count, double p, uint64_t seed_lo,
                                          uint64_t seed_hi, int64_t col_lo, int64_t col_hi, spice_adjacency** out);
/* adj_list::generate(offsets, neighbors, seed)                        (topology.cpp:56-71; bench/connectivity.cpp:8-22)
 * on the GPU: the (src, dst) pairs adj_list::connect collected (host arrays) are sorted by (src, dst) and streamed into
 * CSR.  Same handle, same accessors as above.  An index out of range is a violated precondition (topology.cpp:16-18). */
SPICE_API int spice_adj_list_generate(int device, int32_t const* edges_src, int32_t const* edges_dst, int64_t n_edges, int64_t src_count,
                            int64_t dst_count, int64_t col_lo, int64_t col_hi, spice_adjacency** out);
SPICE_API int64_t spice_adjacency_edges(spice_adjacency const* a);
SPICE_API void* spice_adjacency_offsets_dev(spice_adjacency const* a);   /* int64[src_count+1] */
SPICE_API void* spice_adjacency_neighbors_dev(spice_adjacency const* a); /* int32[edges] */
SPICE_API int spice_adjacency_copy(spice_adjacency const* a, int64_t* offsets_host, int32_t* neighbors_host);
/* entries [edge_lo, edge_hi) of the neighbors array (a block of rows of an adjacency too large to copy whole) */
SPICE_API int spice_adjacency_copy_range(spice_adjacency const* a, int64_t edge_lo, int64_t edge_hi, int32_t* neighbors_host);
/* device-side timings of the last generation, milliseconds: total and the row-writing kernel */
SPICE_API int spice_adjacency_timing(spice_adjacency const* a, float* total_ms, float* rows_kernel_ms, int64_t* draws);
SPICE_API int spice_adjacency_destroy(spice_adjacency* a);

/* ---- seeds (host; random.h:143-175) ---------------------------------------------------------- */
/* snn::_seed (snn.h:69): the seed the next `seed++` hands out -- the one the next connection's
 * Topology::generate receives (synapse_population.h:31, csr.h:69-77). */
SPICE_API int spice_ctx_seed(spice_ctx const* ctx, uint64_t out[2]);
SPICE_API void spice_seed_seq(uint32_t const* words, int n, uint64_t out[2]);
SPICE_API void spice_seed_next(uint64_t seed[2]);
/* FNV-1a (64 bit) over host bytes: the digest of the adjacency golden vectors (tests/golden/golden.json;
 * the reference's own bench/connectivity.cpp has no checksum -- this is the harness's). */
SPICE_API uint64_t spice_fnv1a64(void const* data, int64_t bytes);

/* ---- model tables of the built-in sample models ---------------------------------------------
 * names: "brunel.poisson", "brunel.lif", "vogels.lif"    (samples/brunel.cpp:23-62, vogels.cpp:10-47)
 *        "brunel.fixed_weight", "vogels.excitatory", "vogels.inhibitory", "brunel+.plastic"
 *                                            (samples/brunel.cpp:66-70, vogels.cpp:49-59, brunel+.cpp:59-99) */
SPICE_API spice_neuron_ops const* spice_builtin_neuron(char const* name);
SPICE_API spice_synapse_ops const* spice_builtin_synapse(char const* name);

/* Diagnostics: evaluate on the DEVICE the libm restatements user models reach through
 * spice/util/math.h, so tests can pin them against the host libm the reference links:
 * kind 0: y[i] = exp((float)x[i]) (x, y float32); kind 1: y[i] = pow(x[i], n[i]) (x, y float64, n int64). */
SPICE_API int spice_selftest_libm(int device, int kind, void const* x, int64_t const* n, void* y, int64_t count);

/* device probe: 0 when a CUDA device with compute capability 10.x is usable */
SPICE_API int spice_device_check(int device);
SPICE_API char const* spice_version(void);

#ifdef __cplusplus
}
#endif
#endif
