/* TEST INFRASTRUCTURE — not product code.
 *
 * CPU restatement (plain C11) of the reference's hot path: seeding, xoroshiro128+,
 * fixed_probability adjacency generation, neuron update + spike ring, spike delivery and the
 * snn::step() sequencing, for the sample models (Brunel, Vogels, Brunel+).  Every function in
 * spice_oracle.c cites the reference file:line it follows (paths relative to denniskb/spice2
 * @ f5e57eb).  Parity of this restatement is PINNED: tests/test_oracle.py checks it against the
 * golden vectors in tests/golden/ (generated from the compiled reference, oracle/_ref, by
 * tests/golden/make_golden.py) and, when oracle/_ref is present, directly against the compiled
 * reference on the same seeds.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library.  The product (spice2_b200/) never does.
 */
#ifndef SPICE_ORACLE_H
#define SPICE_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- seeding and RNG ------------------------------------------------------------------- */
typedef struct {
	uint64_t lo, hi;
} orc_u128;

orc_u128 orc_seed_seq(uint32_t const* il, int n);     /* random.h:149-152 */
orc_u128 orc_seed_next(orc_u128 s);                   /* random.h:163-167 (seed++) */
orc_u128 orc_seed_stream(orc_u128 s, uint64_t id);    /* random.h:169 */
void orc_xoroshiro(orc_u128 seed, int64_t count, uint64_t* out); /* random.h:222-234 */
void orc_xoroshiro_state_at(orc_u128 seed, int64_t k, uint64_t state[2]);
/* the same by GF(2) matrix powers, for positions no test can walk to (1e11 draws at 1e6 x 1e6) */
void orc_xoroshiro_jump(orc_u128 seed, uint64_t k, uint64_t state[2]);
void orc_kahan_dt(float dt, int64_t steps, float* out);          /* numeric.h:9-15, snn.cpp:8-10 */
uint64_t orc_fnv1a64(void const* data, int64_t bytes);

/* ---- fixed_probability ------------------------------------------------------------------- */
int64_t orc_fixed_probability_max_degree(int64_t dst, double p); /* topology.cpp:75-78 */
int64_t orc_fixed_probability_size(int64_t src, int64_t dst, double p);
/* topology.cpp:80-112.  offsets has src+1 entries.  neighbors may be NULL (count only).  If
 * row_hash is non-NULL it receives src FNV-1a64 hashes, one per row (streaming check at sizes
 * whose adjacency does not fit memory).  draws_out (optional) = stream position at the end. */
int64_t orc_fixed_probability_generate(int64_t src, int64_t dst, double p, orc_u128 seed,
                                       int64_t* offsets, int32_t* neighbors, uint64_t* row_hash,
                                       int64_t* draws_out);

/* rows of the same loop from the engine state at a row's first draw; targets in [col_lo, col_hi) kept as local columns */
int64_t orc_fixed_probability_rows_from(uint64_t const state[2], int64_t rows, int64_t dst, double p, int64_t col_lo,
                                        int64_t col_hi, int64_t* kept_degree, int64_t* full_degree, int32_t* neighbors,
                                        int64_t cap);

/* ---- networks ---------------------------------------------------------------------------- */
enum { ORC_POISSON = 0, ORC_LIF_BRUNEL = 1, ORC_LIF_VOGELS = 2 };
enum { ORC_FIXED_WEIGHT_V = 0, ORC_WEIGHT_GEX = 1, ORC_WEIGHT_GIN = 2, ORC_PLASTIC_BRUNEL = 3 };
/* floating-point flavour of the neuron/synapse functors:
 *   ORC_STRICT   = the functor's source expression tree evaluated in IEEE RN, no contraction
 *                  (what -fno-fast-math -ffp-contract=off gives; what the CUDA path computes);
 *   ORC_REFBUILD = the forms g++ 13.3 emits under the reference's own flags
 *                  (-O2 -ffast-math -march=haswell), restated from the disassembly. Brunel only. */
enum { ORC_STRICT = 0, ORC_REFBUILD = 1 };

typedef struct orc_net orc_net;

orc_net* orc_net_create(float dt, float max_delay, uint32_t const* seed_il, int n, int flavour);
void orc_net_destroy(orc_net*);
/* target-partitioned execution: this instance owns neurons [size*rank/world, size*(rank+1)/world)
 * of every population.  Must be called before populations are added. */
void orc_net_set_shard(orc_net*, int rank, int world);
int orc_add_population(orc_net*, int model, int64_t size);        /* snn.h:21-27 */
/* snn.h:29-48.  params: [0] = weight for the static synapses. Returns 0, or <0 on a violated
 * precondition (1 <= round(delay/dt) <= max_delay). */
int orc_connect(orc_net*, int src, int dst, double p, float delay, int syn_model, float weight);
void orc_step(orc_net*);                                          /* snn.cpp:7-28 */
/* The same step in two halves, for sharded runs: update local neurons (local spikes are then
 * readable through orc_spikes), install the all-gathered global spike list of each population
 * for this step, then deliver. */
void orc_step_update(orc_net*);
void orc_step_set_spikes(orc_net*, int pop, int32_t const* ids, int64_t n);
void orc_step_deliver(orc_net*);

int64_t orc_population_size(orc_net const*, int pop);
int64_t orc_population_lo(orc_net const*, int pop);
int64_t orc_population_hi(orc_net const*, int pop);
int64_t orc_spikes(orc_net const*, int pop, int64_t age, int32_t const** ids); /* neuron_population.h:147-153 */
/* spikes the local range emitted in the step whose update half just ran (global ids) */
int64_t orc_local_spikes(orc_net const*, int pop, int32_t const** ids);
void const* orc_neurons(orc_net const*, int pop, int64_t* bytes_per_neuron);   /* local range only */
int64_t orc_synaptic_events(orc_net const*);  /* total Syn::deliver calls so far */
int64_t orc_connection_edges(orc_net const*, int conn);
int64_t const* orc_connection_offsets(orc_net const*, int conn);
int32_t const* orc_connection_neighbors(orc_net const*, int conn);
void const* orc_connection_synapses(orc_net const*, int conn, int64_t* bytes_per_synapse);

#ifdef __cplusplus
}
#endif
#endif
