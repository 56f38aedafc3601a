/* TEST INFRASTRUCTURE — not product code.  See spice_oracle.h.
 *
 * CPU restatement of the reference hot path (denniskb/spice2 @ f5e57eb).  Compile with
 * -fno-fast-math -ffp-contract=off: every floating-point operation below is meant literally
 * (IEEE-754 RN); where the compiled reference uses a fused multiply-add the restatement calls
 * fma()/fmaf() explicitly.
 *
 * Third-party arithmetic: the reference calls glibc libm (`log` at random.h:271; `expf`, `pow`
 * at samples/brunel+.cpp:78-79,96-97).  This oracle calls the same libm functions of the host
 * it runs on (Ubuntu GLIBC 2.39-0ubuntu8.5 in this image), exactly as the reference does.
 */
#include "spice_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* murmur3 x64 128 with the reference's custom seed (random.h:32-139)                          */
/* ------------------------------------------------------------------------------------------ */
static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

static inline uint64_t fmix64(uint64_t k) { /* random.h:23-30 */
	k ^= k >> 33;
	k *= 0xff51afd7ed558ccdULL;
	k ^= k >> 33;
	k *= 0xc4ceb9fe1a85ec53ULL;
	k ^= k >> 33;
	return k;
}

#define MM_C1 0x87c37b91114253d5ULL
#define MM_C2 0x4cf5ad432745937fULL
#define MM_H_LO 0x2E4016967F18E81ULL
#define MM_H_HI 0x447567949F9AA86ULL

static orc_u128 mm_finish(orc_u128 h, uint64_t len) { /* random.h:98-107 */
	h.lo ^= len;
	h.hi ^= len;
	h.lo += h.hi;
	h.hi += h.lo;
	h.lo = fmix64(h.lo);
	h.hi = fmix64(h.hi);
	h.lo += h.hi;
	h.hi += h.lo;
	return h;
}

static orc_u128 murmur3_bytes(uint8_t const* data, uint64_t len) { /* random.h:32-107 */
	orc_u128 h       = {MM_H_LO, MM_H_HI};
	uint64_t nblocks = len / 16;
	for (uint64_t i = 0; i < nblocks; i++) {
		uint64_t k1, k2;
		memcpy(&k1, data + i * 16, 8);
		memcpy(&k2, data + i * 16 + 8, 8);
		k1 *= MM_C1;
		k1 = rotl64(k1, 31);
		k1 *= MM_C2;
		h.lo ^= k1;
		h.lo = rotl64(h.lo, 27);
		h.lo += h.hi;
		h.lo = h.lo * 5 + 0x52dce729;
		k2 *= MM_C2;
		k2 = rotl64(k2, 33);
		k2 *= MM_C1;
		h.hi ^= k2;
		h.hi = rotl64(h.hi, 31);
		h.hi += h.lo;
		h.hi = h.hi * 5 + 0x38495ab5;
	}
	uint8_t const* tail = data + nblocks * 16;
	uint64_t k1 = 0, k2 = 0;
	uint64_t rem = len & 15;
	/* random.h:68-96 — the fall-through switch, written as two guarded byte loops */
	if (rem > 8) {
		for (uint64_t b = rem; b > 8; b--)
			k2 ^= (uint64_t)tail[b - 1] << (8 * (b - 9));
		k2 *= MM_C2;
		k2 = rotl64(k2, 33);
		k2 *= MM_C1;
		h.hi ^= k2;
	}
	if (rem > 0) {
		uint64_t top = rem > 8 ? 8 : rem;
		for (uint64_t b = top; b > 0; b--)
			k1 ^= (uint64_t)tail[b - 1] << (8 * (b - 1));
		k1 *= MM_C1;
		k1 = rotl64(k1, 31);
		k1 *= MM_C2;
		h.lo ^= k1;
	}
	return mm_finish(h, len);
}

static orc_u128 murmur3_u128(orc_u128 k) { /* random.h:109-139 */
	orc_u128 h = {MM_H_LO, MM_H_HI};
	k.lo *= MM_C1;
	k.lo = rotl64(k.lo, 31);
	k.lo *= MM_C2;
	h.lo ^= k.lo;
	h.lo = rotl64(h.lo, 27);
	h.lo += h.hi;
	h.lo = h.lo * 5 + 0x52dce729;
	k.hi *= MM_C2;
	k.hi = rotl64(k.hi, 33);
	k.hi *= MM_C1;
	h.hi ^= k.hi;
	h.hi = rotl64(h.hi, 31);
	h.hi += h.lo;
	h.hi = h.hi * 5 + 0x38495ab5;
	return mm_finish(h, 16);
}

orc_u128 orc_seed_seq(uint32_t const* il, int n) { /* random.h:149-152 */
	return murmur3_bytes((uint8_t const*)il, 4u * (uint64_t)n);
}
orc_u128 orc_seed_next(orc_u128 s) { return murmur3_u128(s); } /* random.h:163-167 */
orc_u128 orc_seed_stream(orc_u128 s, uint64_t id) {            /* random.h:169, stdint.h:19-23 */
	uint64_t n = id + 1;
	orc_u128 r = {s.lo + n, s.hi};
	r.hi += r.lo < s.lo;
	return murmur3_u128(r);
}

/* ------------------------------------------------------------------------------------------ */
/* xoroshiro128+ (random.h:222-234), state = (seed.lo, seed.hi) (random.h:182,188-192)          */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
	uint64_t s0, s1;
} xoro;

static inline xoro xoro_init(orc_u128 seed) {
	xoro r = {seed.lo, seed.hi};
	return r;
}
static inline uint64_t xoro_next(xoro* r) {
	uint64_t const result = r->s0 + r->s1;
	uint64_t const tmp    = r->s0 ^ r->s1;
	r->s0                 = rotl64(r->s0, 24) ^ tmp ^ (tmp << 16);
	r->s1                 = rotl64(tmp, 37);
	return result;
}

void orc_xoroshiro(orc_u128 seed, int64_t count, uint64_t* out) {
	xoro r = xoro_init(seed);
	for (int64_t i = 0; i < count; i++)
		out[i] = xoro_next(&r);
}

void orc_xoroshiro_state_at(orc_u128 seed, int64_t k, uint64_t state[2]) {
	xoro r = xoro_init(seed);
	for (int64_t i = 0; i < k; i++)
		(void)xoro_next(&r);
	state[0] = r.s0;
	state[1] = r.s1;
}

/* Engine state after k draws without walking them.  The state update of xoroshiro128+ (random.h:222-234) is linear over
 * GF(2): one step is a 128 x 128 bit matrix T, k steps are T^k, built from T^(2^i) by repeated squaring.  For streams no
 * CPU can walk in a test (bench/connectivity at 1e6 x 1e6 draws 1e11 values); checked against the walk in tests. */
typedef struct {
	uint64_t lo, hi;
} bits128;
static bits128 mat_vec(bits128 const m[128], bits128 v) {
	bits128 r = {0, 0};
	for (int j = 0; j < 128; j++)
		if ((j < 64 ? v.lo >> j : v.hi >> (j - 64)) & 1u) {
			r.lo ^= m[j].lo;
			r.hi ^= m[j].hi;
		}
	return r;
}
void orc_xoroshiro_jump(orc_u128 seed, uint64_t k, uint64_t state[2]) {
	bits128 m[128], sq[128];
	for (int j = 0; j < 128; j++) { /* column j of T: one step from the basis state e_j */
		xoro r = {j < 64 ? 1ull << j : 0, j < 64 ? 0 : 1ull << (j - 64)};
		(void)xoro_next(&r);
		m[j].lo = r.s0;
		m[j].hi = r.s1;
	}
	bits128 v = {seed.lo, seed.hi};
	for (; k; k >>= 1) {
		if (k & 1)
			v = mat_vec(m, v);
		for (int j = 0; j < 128; j++)
			sq[j] = mat_vec(m, m[j]);
		memcpy(m, sq, sizeof m);
	}
	state[0] = v.lo;
	state[1] = v.hi;
}

/* generate_canonical<float,false> on a 64-bit engine (random.h:236-247): top 24 bits / 2^24 */
static inline float canonical_float(xoro* r) { return (float)(xoro_next(r) >> 40) / 16777216.0f; }
/* generate_canonical<double,true>: ((r >> 11) + 1) / 2^53, in (0,1] */
static inline double canonical_double_leftopen(xoro* r) {
	return (double)((xoro_next(r) >> 11) + 1) / 9007199254740992.0;
}

/* ------------------------------------------------------------------------------------------ */
/* kahan-compensated dt (numeric.h:9-15 under its -fno-fast-math attribute; snn.cpp:8-10)       */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
	float c, sum;
} kahan;

static inline float kahan_add(kahan* k, float delta) {
	float const y = delta - k->c;
	float const t = k->sum + y;
	k->c          = (t - k->sum) - y;
	k->sum        = t;
	return y;
}

void orc_kahan_dt(float dt, int64_t steps, float* out) {
	kahan k = {0, 0};
	for (int64_t i = 0; i < steps; i++) {
		out[i] = kahan_add(&k, dt);
		if (k.sum >= 1)
			k.sum = 0; /* reset() clears the sum only, numeric.h:19 */
	}
}

uint64_t orc_fnv1a64(void const* data, int64_t bytes) {
	uint8_t const* p = (uint8_t const*)data;
	uint64_t h       = 0xcbf29ce484222325ULL;
	for (int64_t i = 0; i < bytes; i++)
		h = (h ^ p[i]) * 0x100000001b3ULL;
	return h;
}

/* ------------------------------------------------------------------------------------------ */
/* fixed_probability (topology.cpp:75-112), in the floating-point forms g++ 13.3 emits under    */
/* the reference's flags (verified by disassembly of oracle/_ref/topology_fast.o):              */
/*   max_degree = (int64) trunc( fma( sqrt((1-p) * (dst*p)), 3.0, dst*p ) )                     */
/*   c          = 1.0 - (1.0 / p)                      [= -scale, scale = 1/p - 1]              */
/*   u          = (double)((r >> 11) + 1) * 2^-53                                               */
/*   noise      = fma(log(u), c, noise)                                                         */
/*   dst        = index + cvttsd2si32( noise + copysign(0x1.fffffffffffffp-2, noise) )          */
/* ------------------------------------------------------------------------------------------ */
int64_t orc_fixed_probability_max_degree(int64_t dst, double p) {
	double const dp = (double)dst * p;
	double const v  = (1.0 - p) * dp;
	return (int64_t)fma(sqrt(v), 3.0, dp);
}

int64_t orc_fixed_probability_size(int64_t src, int64_t dst, double p) {
	return src * orc_fixed_probability_max_degree(dst, p);
}

/* x86 cvttsd2si (32-bit): out-of-range and NaN give INT32_MIN */
static inline int32_t cvttsd2si32(double x) {
	if (!(x > -2147483649.0 && x < 2147483648.0))
		return INT32_MIN;
	return (int32_t)x;
}

int64_t orc_fixed_probability_generate(int64_t src, int64_t dst, double p, orc_u128 seed,
                                       int64_t* offsets, int32_t* neighbors, uint64_t* row_hash,
                                       int64_t* draws_out) {
	if (draws_out)
		*draws_out = 0;
	if (src == 0 || dst == 0 || p == 0) /* topology.cpp:85-86: arrays are left untouched */
		return 0;

	xoro rng                 = xoro_init(seed);
	double const c           = 1.0 - 1.0 / p;
	int64_t const max_degree = orc_fixed_probability_max_degree(dst, p);
	int64_t count            = 0;
	int64_t draws            = 0;
	for (int64_t s = 0; s < src; s++) {
		if (offsets)
			offsets[s] = count;
		int32_t index = 0;
		double noise  = 0;
		uint64_t h    = 0xcbf29ce484222325ULL;
		for (;;) {
			double const u = (double)((xoro_next(&rng) >> 11) + 1) * 0x1p-53;
			draws++;
			noise             = fma(log(u), c, noise);
			int32_t const d32 = (int32_t)((uint32_t)index +
			                              (uint32_t)cvttsd2si32(noise + copysign(0x1.fffffffffffffp-2, noise)));
			if (((int64_t)d32 >= dst) | (index >= max_degree))
				break;
			if (neighbors)
				neighbors[count] = d32;
			if (row_hash)
				for (int b = 0; b < 4; b++)
					h = (h ^ (((uint32_t)d32 >> (8 * b)) & 0xff)) * 0x100000001b3ULL;
			count++;
			index++;
		}
		if (row_hash)
			row_hash[s] = h;
	}
	if (offsets)
		offsets[src] = count;
	if (draws_out)
		*draws_out = draws;
	return count;
}

/* `rows` consecutive rows of the same loop (topology.cpp:89-110), started from the engine state at the first row's first
 * draw: for matrices whose stream is entered by orc_xoroshiro_jump.  Only targets in [col_lo, col_hi) are kept, as local
 * columns (a rank's share of a sharded adjacency); kept_degree / full_degree receive one entry per row.  Returns the
 * number of kept entries (those beyond `cap` are counted, not written). */
int64_t orc_fixed_probability_rows_from(uint64_t const state[2], int64_t rows, int64_t dst, double p, int64_t col_lo,
                                        int64_t col_hi, int64_t* kept_degree, int64_t* full_degree, int32_t* neighbors,
                                        int64_t cap) {
	xoro rng                 = {state[0], state[1]};
	double const c           = 1.0 - 1.0 / p;
	int64_t const max_degree = orc_fixed_probability_max_degree(dst, p);
	int64_t count            = 0;
	for (int64_t s = 0; s < rows; s++) {
		int32_t index = 0;
		int64_t kept  = 0;
		double noise  = 0;
		for (;;) {
			double const u    = (double)((xoro_next(&rng) >> 11) + 1) * 0x1p-53;
			noise             = fma(log(u), c, noise);
			int32_t const d32 = (int32_t)((uint32_t)index +
			                              (uint32_t)cvttsd2si32(noise + copysign(0x1.fffffffffffffp-2, noise)));
			if (((int64_t)d32 >= dst) | (index >= max_degree))
				break;
			if (d32 >= col_lo && d32 < col_hi) {
				if (neighbors && count < cap)
					neighbors[count] = (int32_t)(d32 - col_lo);
				count++;
				kept++;
			}
			index++;
		}
		if (kept_degree)
			kept_degree[s] = kept;
		if (full_degree)
			full_degree[s] = index;
	}
	return count;
}

/* ------------------------------------------------------------------------------------------ */
/* models: samples/brunel.cpp:23-72, samples/vogels.cpp:10-59, samples/brunel+.cpp:59-99        */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
	float V;
	int32_t Twait;
} lif_brunel; /* brunel.cpp:40-43 */

typedef struct {
	float V, Gex, Gin;
	int32_t Twait;
} lif_vogels; /* vogels.cpp:11-16 */

typedef struct {
	float W, Zpre, Zpost;
} syn_plastic; /* brunel+.cpp:62-66 */

/* brunel.cpp:27-30 */
static inline int poisson_update(float dt, xoro* rng) {
	float const firing_rate = 20;
	return canonical_float(rng) < (firing_rate * dt);
}

/* brunel.cpp:45-61, strict: V += (Vrest - V) * (dt * TmemInv) */
static inline int lif_brunel_update_strict(lif_brunel* n, float dt) {
	float const TmemInv = (float)(1.0 / 0.02);
	float const Vrest   = 0.0f;
	int const Tref      = 20;
	float const Vthres  = (float)0.02;
	if (--n->Twait <= 0) {
		if (n->V > Vthres) {
			n->V     = Vrest;
			n->Twait = Tref;
			return 1;
		}
		float const a = Vrest - n->V;
		float const b = dt * TmemInv;
		float const m = a * b;
		n->V          = n->V + m;
	}
	return 0;
}

/* the same functor as g++ -O2 -ffast-math -march=haswell compiles it inside
 * neuron_population<lif>::update (disassembly of oracle/_ref/brunel): k = fma(dt,-50,1) hoisted
 * out of the loop; V *= k */
static inline int lif_brunel_update_refbuild(lif_brunel* n, float k) {
	if (--n->Twait <= 0) {
		if (n->V > (float)0.02) {
			n->V     = 0.0f;
			n->Twait = 20;
			return 1;
		}
		n->V = n->V * k;
	}
	return 0;
}

/* vogels.cpp:18-46, strict */
static inline int lif_vogels_update_strict(lif_vogels* n, float dt) {
	int32_t const Tref  = 50;
	float const Vrest   = (float)-0.06;
	float const Vthres  = (float)-0.05;
	float const TmemInv = (float)(1.0f / 0.02);
	float const Eex     = (float)0.0;
	float const Ein     = (float)-0.08;
	float const Ibg     = (float)0.02;
	float const TexInv  = (float)(1.0f / 0.005);
	float const TinInv  = (float)(1.0f / 0.01);

	int spiked = 0;
	if (--n->Twait <= 0) {
		if (n->V > Vthres) {
			n->V     = Vrest;
			n->Twait = Tref;
			spiked   = 1;
		} else {
			float const t0 = Vrest - n->V;
			float const t1 = n->Gex * (Eex - n->V);
			float const t2 = n->Gin * (Ein - n->V);
			float const s  = ((t0 + t1) + t2) + Ibg;
			n->V           = n->V + s * (dt * TmemInv);
		}
	}
	n->Gex = n->Gex - n->Gex * (dt * TexInv);
	n->Gin = n->Gin - n->Gin * (dt * TinInv);
	return spiked;
}

/* brunel+.cpp:74-89 */
static inline float clampf(float v, float lo, float hi) { /* std::clamp */
	return (v < lo) ? lo : (hi < v) ? hi : v;
}
static inline void plastic_update(syn_plastic* s, float dt, int pre, int post) {
	float const TstdpInv = 1.0f / 0.02f;
	float const dtInv    = 1.0f / dt;
	float const fpre     = (float)pre;
	float const fpost    = (float)post;
	float const a        = ((fpre * 0.0202f) * s->W) * expf(-s->Zpost * dtInv);
	float const b        = ((fpost * 0.01f) * (1.0f - s->W)) * expf(-s->Zpre * dtInv);
	s->W                 = clampf((s->W - a) + b, 0.0f, 0.0003f);
	s->Zpre              = s->Zpre + fpre;
	s->Zpost             = s->Zpost + fpost;
	s->Zpre              = s->Zpre - (s->Zpre * dt) * TstdpInv;
	s->Zpost             = s->Zpost - (s->Zpost * dt) * TstdpInv;
}
/* brunel+.cpp:93-98: std::pow(float, Int) promotes to double pow; float *= double */
static inline void plastic_skip(syn_plastic* s, float dt, int64_t n) {
	float const TstdpInv = 1.0f / 0.02f;
	float const base     = 1 - dt * TstdpInv;
	double const f       = pow((double)base, (double)n);
	s->Zpre              = (float)((double)s->Zpre * f);
	s->Zpost             = (float)((double)s->Zpost * f);
}

/* ------------------------------------------------------------------------------------------ */
/* network                                                                                      */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
	int model;
	int64_t size, lo, hi; /* global size; local range [lo,hi) */
	void* neurons;        /* local AoS state (NULL for stateless) */
	/* global spike ring (neuron_population.h:119-122,133,147-153) */
	int32_t* spikes;
	int64_t spikes_len, spikes_cap;
	int32_t* counts;
	int64_t counts_len;
	/* spikes emitted by the local range in the step being run */
	int32_t* local;
	int64_t local_len, local_cap;
	int plastic;
	uint64_t* history; /* local range, neuron_population.h:126-132,155-160 */
} population;

typedef struct {
	int src, dst, model;
	int64_t delay;
	float weight;
	int64_t* offsets;
	int32_t* neighbors;
	int64_t edges;
	syn_plastic* syn; /* parallel to neighbors (csr.h:97-99) */
	uint64_t* ages;   /* per source (synapse_population.h:80,43-44) */
} connection;

struct orc_net {
	float dt;
	int64_t max_delay;
	int64_t time;
	kahan simtime;
	orc_u128 seed;
	int flavour;
	int rank, world;
	population* pops;
	int npops;
	connection* conns;
	int nconns;
	int64_t events;
	float step_dt; /* compensated dt of the step in flight */
};

orc_net* orc_net_create(float dt, float max_delay, uint32_t const* il, int n, int flavour) {
	orc_net* net   = (orc_net*)calloc(1, sizeof(orc_net));
	net->dt        = dt;
	net->max_delay = (int64_t)roundf(max_delay / dt); /* snn.h:18-19 */
	net->seed      = orc_seed_seq(il, n);
	net->flavour   = flavour;
	net->world     = 1;
	return net;
}

void orc_net_destroy(orc_net* net) {
	if (!net)
		return;
	for (int i = 0; i < net->npops; i++) {
		free(net->pops[i].neurons);
		free(net->pops[i].spikes);
		free(net->pops[i].counts);
		free(net->pops[i].local);
		free(net->pops[i].history);
	}
	for (int i = 0; i < net->nconns; i++) {
		free(net->conns[i].offsets);
		free(net->conns[i].neighbors);
		free(net->conns[i].syn);
		free(net->conns[i].ages);
	}
	free(net->pops);
	free(net->conns);
	free(net);
}

void orc_net_set_shard(orc_net* net, int rank, int world) {
	net->rank  = rank;
	net->world = world;
}

static size_t neuron_bytes(int model) {
	return model == ORC_LIF_BRUNEL ? sizeof(lif_brunel) : model == ORC_LIF_VOGELS ? sizeof(lif_vogels) : 0;
}

int orc_add_population(orc_net* net, int model, int64_t size) {
	net->pops     = (population*)realloc(net->pops, sizeof(population) * (size_t)(net->npops + 1));
	population* p = &net->pops[net->npops];
	memset(p, 0, sizeof *p);
	p->model = model;
	p->size  = size;
	p->lo    = size * net->rank / net->world;
	p->hi    = size * (net->rank + 1) / net->world;
	int64_t const n = p->hi - p->lo;
	if (model == ORC_LIF_BRUNEL) {
		lif_brunel* v = (lif_brunel*)calloc((size_t)(n > 0 ? n : 1), sizeof(lif_brunel));
		p->neurons    = v; /* V = 0, Twait = 0 */
	} else if (model == ORC_LIF_VOGELS) {
		lif_vogels* v = (lif_vogels*)calloc((size_t)(n > 0 ? n : 1), sizeof(lif_vogels));
		for (int64_t i = 0; i < n; i++)
			v[i].V = (float)-0.06;
		p->neurons = v;
	}
	/* the stateful adapter burns one seed++ even without init() (neuron_population.h:60);
	 * the stateless adapter burns none (neuron_population.h:32-36) */
	if (model != ORC_POISSON)
		net->seed = orc_seed_next(net->seed);
	p->counts = (int32_t*)calloc((size_t)net->max_delay + 1, sizeof(int32_t));
	return net->npops++;
}

int orc_connect(orc_net* net, int src, int dst, double p, float delay, int syn_model, float weight) {
	int64_t const d = (int64_t)roundf(delay / net->dt); /* snn.h:33 */
	if (!(d >= 1) || !(d <= net->max_delay))            /* snn.h:34-38 */
		return -1;
	net->conns    = (connection*)realloc(net->conns, sizeof(connection) * (size_t)(net->nconns + 1));
	connection* c = &net->conns[net->nconns];
	memset(c, 0, sizeof *c);
	c->src    = src;
	c->dst    = dst;
	c->model  = syn_model;
	c->delay  = d;
	c->weight = weight;
	int64_t const ns  = net->pops[src].size;
	int64_t const nd  = net->pops[dst].size;
	int64_t const cap = orc_fixed_probability_size(ns, nd, p); /* csr.h:70-73 */
	c->offsets        = (int64_t*)calloc((size_t)ns + 1, sizeof(int64_t));
	c->neighbors      = (int32_t*)malloc(sizeof(int32_t) * (size_t)(cap > 0 ? cap : 1));
	/* synapse_population ctor: _graph(c, seed++) (synapse_population.h:30-31) */
	c->edges  = orc_fixed_probability_generate(ns, nd, p, net->seed, c->offsets, c->neighbors, NULL, NULL);
	net->seed = orc_seed_next(net->seed);
	if (syn_model == ORC_PLASTIC_BRUNEL) {
		c->syn = (syn_plastic*)malloc(sizeof(syn_plastic) * (size_t)(c->edges > 0 ? c->edges : 1));
		for (int64_t i = 0; i < c->edges; i++) {
			c->syn[i].W     = (float)1e-4; /* brunel+.cpp:63-65 */
			c->syn[i].Zpre  = 0;
			c->syn[i].Zpost = 0;
		}
		c->ages = (uint64_t*)calloc((size_t)ns, sizeof(uint64_t)); /* synapse_population.h:43-44 */
		/* snn.h:46-47: source->plastic() */
		population* sp = &net->pops[src];
		if (!sp->plastic) {
			sp->plastic = 1;
			sp->history = (uint64_t*)calloc((size_t)(sp->hi - sp->lo > 0 ? sp->hi - sp->lo : 1), sizeof(uint64_t));
		}
	}
	net->nconns++;
	return 0;
}

static void push_i32(int32_t** buf, int64_t* len, int64_t* cap, int32_t v) {
	if (*len == *cap) {
		*cap = *cap ? *cap * 2 : 1024;
		*buf = (int32_t*)realloc(*buf, sizeof(int32_t) * (size_t)*cap);
	}
	(*buf)[(*len)++] = v;
}

/* neuron_population::update (neuron_population.h:116-134), first half: evict + adapter.update */
void orc_step_update(orc_net* net) {
	/* snn.cpp:8-12 */
	net->step_dt = kahan_add(&net->simtime, net->dt);
	if (net->simtime.sum >= 1)
		net->simtime.sum = 0;
	xoro rng  = xoro_init(net->seed);
	net->seed = orc_seed_next(net->seed);
	float const dt = net->step_dt;
	float const k  = fmaf(dt, -50.0f, 1.0f); /* ORC_REFBUILD form of the Brunel leak factor */

	for (int pi = 0; pi < net->npops; pi++) {
		population* p = &net->pops[pi];
		/* neuron_population.h:119-122 */
		if (p->counts_len == net->max_delay) {
			int64_t const front = p->counts[0];
			memmove(p->spikes, p->spikes + front, sizeof(int32_t) * (size_t)(p->spikes_len - front));
			p->spikes_len -= front;
			memmove(p->counts, p->counts + 1, sizeof(int32_t) * (size_t)(p->counts_len - 1));
			p->counts_len--;
		}
		p->local_len = 0;
		/* neuron_population.h:40-45 / 72-77: ascending index; ONE rng shared by all populations
		 * of the step (snn.cpp:12-15).  A shard draws and discards for neurons it does not own. */
		if (p->model == ORC_POISSON) {
			for (int64_t i = 0; i < p->size; i++) {
				int const s = poisson_update(dt, &rng);
				if (s && i >= p->lo && i < p->hi)
					push_i32(&p->local, &p->local_len, &p->local_cap, (int32_t)i);
			}
		} else if (p->model == ORC_LIF_BRUNEL) {
			lif_brunel* n = (lif_brunel*)p->neurons;
			for (int64_t i = p->lo; i < p->hi; i++) {
				int const s = net->flavour == ORC_REFBUILD ? lif_brunel_update_refbuild(&n[i - p->lo], k) :
				                                             lif_brunel_update_strict(&n[i - p->lo], dt);
				if (s)
					push_i32(&p->local, &p->local_len, &p->local_cap, (int32_t)i);
			}
		} else {
			lif_vogels* n = (lif_vogels*)p->neurons;
			for (int64_t i = p->lo; i < p->hi; i++)
				if (lif_vogels_update_strict(&n[i - p->lo], dt))
					push_i32(&p->local, &p->local_len, &p->local_cap, (int32_t)i);
		}
		/* neuron_population.h:126-132 */
		if (p->plastic) {
			for (int64_t i = 0; i < p->hi - p->lo; i++)
				p->history[i] <<= 1;
			for (int64_t i = 0; i < p->local_len; i++)
				p->history[p->local[i] - p->lo] |= 1;
		}
	}
}

/* second half of neuron_population::update: append this step's (global) spike list to the ring */
void orc_step_set_spikes(orc_net* net, int pop, int32_t const* ids, int64_t n) {
	population* p = &net->pops[pop];
	for (int64_t i = 0; i < n; i++)
		push_i32(&p->spikes, &p->spikes_len, &p->spikes_cap, ids[i]);
	p->counts[p->counts_len++] = (int32_t)n; /* neuron_population.h:133 */
}

/* neuron_population::spikes(age) (neuron_population.h:147-153) */
int64_t orc_spikes(orc_net const* net, int pop, int64_t age, int32_t const** ids) {
	population const* p = &net->pops[pop];
	if (!(0 <= age && age < p->counts_len)) /* SPICE_PRE */
		return -1;
	int64_t offset = 0;
	for (int64_t i = p->counts_len - 1 - age; i < p->counts_len; i++)
		offset += p->counts[i];
	if (ids)
		*ids = p->spikes + p->spikes_len - offset;
	return p->counts[p->counts_len - 1 - age];
}

/* synapse_population::_update<Deliver> (synapse_population.h:82-140) for one source */
static void update_source(orc_net* net, connection* c, int deliver, int64_t src) {
	population* dp     = &net->pops[c->dst];
	int64_t const time = net->time;
	float const dt     = net->dt; /* nominal dt, snn.cpp:19,23 */
	int const plastic  = c->model == ORC_PLASTIC_BRUNEL;

	int pre     = 0;
	int64_t age = time + 1;
	if (plastic) { /* :91-94 */
		pre = (int)(c->ages[src] >> 63);
		age = (int64_t)(c->ages[src] & ~(1ULL << 63));
	}
	int64_t const prefix = 63 + pre - time + age;                      /* :95 */
	uint64_t const mask  = prefix < 64 ? (~0ULL >> prefix) : 0;        /* :96 (UB when >= 64; unused then) */
	int const outdated   = time >= age;

	for (int64_t e = c->offsets[src]; e < c->offsets[src + 1]; e++) {
		int64_t const dst = c->neighbors[e];
		if (dst < dp->lo || dst >= dp->hi)
			continue; /* target-partitioned shard: only local targets */
		if (plastic && outdated) { /* :100-116 */
			syn_plastic* s = &c->syn[e];
			uint64_t hist  = dp->history[dst - dp->lo];
			if (pre)
				plastic_update(s, dt, 1, (hist & (1ULL << (time - age))) != 0);
			hist &= mask;
			int64_t p = prefix;
			while (hist) {
				int64_t const lz = __builtin_clzll(hist);
				plastic_skip(s, dt, lz - p);
				plastic_update(s, dt, 0, 1);
				hist ^= 1ULL << (63 - lz);
				p = lz + 1;
			}
			plastic_skip(s, dt, 64 - p);
		}
		if (deliver) { /* :118-133 */
			net->events++;
			switch (c->model) {
				case ORC_FIXED_WEIGHT_V: { /* brunel.cpp:67-70 */
					lif_brunel* n = (lif_brunel*)dp->neurons;
					n[dst - dp->lo].V += c->weight;
				} break;
				case ORC_WEIGHT_GEX: { /* vogels.cpp:49-52 */
					lif_vogels* n = (lif_vogels*)dp->neurons;
					n[dst - dp->lo].Gex += c->weight;
				} break;
				case ORC_WEIGHT_GIN: { /* vogels.cpp:55-58 */
					lif_vogels* n = (lif_vogels*)dp->neurons;
					n[dst - dp->lo].Gin += c->weight;
				} break;
				case ORC_PLASTIC_BRUNEL: { /* brunel+.cpp:70 */
					lif_brunel* n = (lif_brunel*)dp->neurons;
					n[dst - dp->lo].V += c->syn[e].W;
				} break;
			}
		}
	}
	if (plastic) /* :137-138 */
		c->ages[src] = (uint64_t)(time + 1) | ((uint64_t)(deliver != 0) << 63);
}

void orc_step_deliver(orc_net* net) {
	/* snn.cpp:17-19: every 64 steps catch every plastic synapse up (synapse_population.h:68-72) */
	if (net->time % 64 == 0)
		for (int ci = 0; ci < net->nconns; ci++)
			if (net->conns[ci].model == ORC_PLASTIC_BRUNEL)
				for (int64_t s = 0; s < net->pops[net->conns[ci].src].size; s++)
					update_source(net, &net->conns[ci], 0, s);
	/* snn.cpp:21-25 */
	for (int ci = 0; ci < net->nconns; ci++) {
		connection* c = &net->conns[ci];
		if (net->time >= c->delay - 1) {
			int32_t const* ids;
			int64_t const n = orc_spikes(net, c->src, c->delay - 1, &ids);
			for (int64_t i = 0; i < n; i++)
				update_source(net, c, 1, ids[i]);
		}
	}
	net->time++;
}

void orc_step(orc_net* net) {
	orc_step_update(net);
	for (int pi = 0; pi < net->npops; pi++)
		orc_step_set_spikes(net, pi, net->pops[pi].local, net->pops[pi].local_len);
	orc_step_deliver(net);
}

int64_t orc_population_size(orc_net const* net, int pop) { return net->pops[pop].size; }
int64_t orc_population_lo(orc_net const* net, int pop) { return net->pops[pop].lo; }
int64_t orc_population_hi(orc_net const* net, int pop) { return net->pops[pop].hi; }
int64_t orc_local_spikes(orc_net const* net, int pop, int32_t const** ids) {
	if (ids)
		*ids = net->pops[pop].local;
	return net->pops[pop].local_len;
}
void const* orc_neurons(orc_net const* net, int pop, int64_t* bytes_per_neuron) {
	if (bytes_per_neuron)
		*bytes_per_neuron = (int64_t)neuron_bytes(net->pops[pop].model);
	return net->pops[pop].neurons;
}
int64_t orc_synaptic_events(orc_net const* net) { return net->events; }
int64_t orc_connection_edges(orc_net const* net, int conn) { return net->conns[conn].edges; }
int64_t const* orc_connection_offsets(orc_net const* net, int conn) { return net->conns[conn].offsets; }
int32_t const* orc_connection_neighbors(orc_net const* net, int conn) { return net->conns[conn].neighbors; }
void const* orc_connection_synapses(orc_net const* net, int conn, int64_t* bytes_per_synapse) {
	if (bytes_per_synapse)
		*bytes_per_synapse = net->conns[conn].syn ? (int64_t)sizeof(syn_plastic) : 0;
	return net->conns[conn].syn;
}
