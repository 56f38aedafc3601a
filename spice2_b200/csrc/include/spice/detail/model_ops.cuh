// Kernel templates instantiated per user model, and the type-erased tables (spice_neuron_ops /
// spice_synapse_ops) that hand them to the model-agnostic runtime.
//
// What runs here is the reference's per-neuron loop, neuron_population<Neur>::update
// (spice/include/spice/detail/neuron_population.h:116-134) with its adapters (:40-45, :72-77),
// re-designed for the GPU:
//   * neuron state is word-SoA (word w of neuron i at state[w*stride + i]) so every load/store
//     of a warp is one coalesced 128-byte line;
//   * one launch advances a whole WINDOW of steps (<= the minimum synaptic delay): inside a
//     window a neuron depends only on its own state and on event counters written before the
//     window began, so a thread keeps its neuron in registers across all steps;
//   * incoming synaptic events arrive as per-(slot, connection, target) integer counters; the
//     owner thread applies Syn::deliver that many times, connection by connection in connect()
//     order — the same sequence of float operations the reference performs
//     (synapse_population.h:118-133 under snn.cpp:21-25), hence bit-exact;
//   * spikes are compacted with warp ballots into the step's spike list (and stored straight
//     into every peer GPU's copy of the list over NVLink when world > 1);
//   * the step's single xoroshiro stream (snn.cpp:12-15) is entered at arbitrary offsets through
//     GF(2) jump polynomials (spice/util/random.h), 32 consecutive neurons per jump.
//
// Everything in this header must be compiled into ONE device module together with the user's
// functors (device function pointers are only valid inside the module that defines them).
// Compile with -fmad=false so the functors' float expressions are evaluated as written.
#pragma once

#include <cuda_runtime.h>

#include <cstring>
#include <new>
#include <span>
#include <type_traits>

#include "spice/concepts.h"
#include "spice/detail/abi.h"
#include "spice/util/random.h"

namespace spice::detail {

// ---- small helpers ------------------------------------------------------------------------------
template <class T>
struct words_of {
	static_assert(std::is_trivially_copyable_v<T>, "neuron / synapse state must be trivially copyable");
	static constexpr int value = (sizeof(T) + 3) / 4; // the last word of a state whose size is no multiple of 4 is padded
};

template <class T>
__device__ __forceinline__ T load_soa(std::uint32_t const* base, std::int64_t stride, std::int64_t i) {
	constexpr int W = words_of<T>::value;
	std::uint32_t w[W];
#pragma unroll
	for (int k = 0; k < W; k++)
		w[k] = base[k * stride + i];
	T out;
	memcpy(&out, w, sizeof(T));
	return out;
}

template <class T>
__device__ __forceinline__ void store_soa(std::uint32_t* base, std::int64_t stride, std::int64_t i, T const& v) {
	constexpr int W = words_of<T>::value;
	std::uint32_t w[W] = {};
	memcpy(w, &v, sizeof(T));
#pragma unroll
	for (int k = 0; k < W; k++)
		base[k * stride + i] = w[k];
}

// engine handed to models that declare no draws: using it is a compile-time error
struct null_rng {
	using result_type = UInt;
	template <class T = void>
	__host__ __device__ UInt operator()() {
		static_assert(!std::is_void_v<T>,
		              "this neuron draws from the rng: declare `static constexpr int rng_draws = <draws per update()>;`");
		return 0;
	}
};

// engine that counts its draws so a model that uses fewer than it declared stays in step
struct counting_rng {
	using result_type = UInt;
	util::xoroshiro64_128p g;
	int used = 0;
	__device__ __forceinline__ UInt operator()() {
		used++;
		return g();
	}
};

// ---- apply: k deliveries of a stateless synapse to one neuron ------------------------------------
// k sequential deliveries on a register copy of the neuron: the remainder first (predicated), then
// groups of 8 back to back — a dependent chain of the functor's own float operations, nothing else
template <class Syn, class N>
__device__ __forceinline__ void deliver_k(Syn const& syn, N& n, unsigned k) {
#pragma unroll
	for (unsigned u = 1; u < 8; u++)
		if (u <= (k & 7u))
			syn.deliver(n);
	for (unsigned j = k >> 3; j; j--) {
#pragma unroll
		for (int u = 0; u < 8; u++)
			syn.deliver(n);
	}
}

template <class Syn, class DstNeur>
__device__ void apply_impl(void const* functor, void* neuron, unsigned k) {
	if constexpr (!StatefulSynapse<Syn> && DeliverTo<Syn, DstNeur>) {
		using N       = typename DstNeur::neuron;
		Syn const syn = *static_cast<Syn const*>(functor);
		N n           = *static_cast<N*>(neuron);
		deliver_k(syn, n, k);
		*static_cast<N*>(neuron) = n;
	}
}
template <class Syn, class DstNeur>
__device__ apply_fn apply_ptr = apply_impl<Syn, DstNeur>;

// ---- stateful / plastic synapses -------------------------------------------------------------------
// apply_events: the events one stateful connection addressed to one neuron in the step that just ran
template <class Syn, class SrcNeur, class DstNeur>
__device__ void apply_events_impl(void const* functor, void* neuron, std::uint32_t const* syn, std::int64_t syn_stride,
                                  std::int32_t* list, unsigned n, from_ctx const* from) {
	if constexpr (StatefulSynapse<Syn>) {
		using N = typename DstNeur::neuron;
		using S = typename Syn::synapse;
		// the step's events of this target, sorted: ascending edge index = (source, row) order.  Lists of up to kLocal
		// events are sorted in a thread-local copy, not in global memory; Shell sort (Ciura's gaps): a target of
		// samples/brunel+ receives ~20 events in a quiet step and a few hundred in a population burst, where an insertion
		// sort's n^2 / 4 moves made this the longest kernel of the step
		constexpr unsigned kLocal = 256;
		std::int32_t local[kLocal];
		if (n <= kLocal) {
			for (unsigned i = 0; i < n; i++)
				local[i] = list[i];
			list = local;
		}
		if (!from->unordered) {
			constexpr unsigned gaps[8] = {701, 301, 132, 57, 23, 10, 4, 1};
			for (unsigned gi = 0; gi < 8; gi++) {
				unsigned const gap = gaps[gi];
				for (unsigned i = gap; i < n; i++) {
					std::int32_t const key = list[i];
					unsigned j             = i;
					for (; j >= gap && list[j - gap] > key; j -= gap)
						list[j] = list[j - gap];
					list[j] = key;
				}
			}
		}
		Syn const f = *static_cast<Syn const*>(functor);
		N nn        = *static_cast<N*>(neuron);
		for (unsigned i = 0; i < n; i++) {
			S const sy = load_soa<S>(syn, syn_stride, list[i]);
			if constexpr (DeliverTo<Syn, DstNeur>)
				f.deliver(sy, nn);
			else if constexpr (detail::deliver_from_to_v<Syn, SrcNeur, DstNeur>) {
				// the row that holds edge list[i]: last src with offsets[src] <= e
				std::int64_t lo = 0, hi = from->n_src;
				while (hi - lo > 1) {
					std::int64_t const mid = (lo + hi) >> 1;
					if (from->offsets[mid] <= list[i])
						lo = mid;
					else
						hi = mid;
				}
				auto const sn = load_soa<typename SrcNeur::neuron>(from->state, from->stride, lo);
				f.deliver(sy, sn, nn);
			}
		}
		*static_cast<N*>(neuron) = nn;
	}
	(void)functor, (void)neuron, (void)syn, (void)syn_stride, (void)list, (void)n, (void)from;
}
template <class Syn, class SrcNeur, class DstNeur>
__device__ apply_events_fn apply_events_ptr = apply_events_impl<Syn, SrcNeur, DstNeur>;

// Lazy plasticity: bring one synapse from step `age` up to and including step `time`
// (synapse_population.h:95-116 with Outdated == true).  hist bit j = the target fired at step time - j.
template <class Syn>
__device__ __forceinline__ void catch_up(Syn const& f, typename Syn::synapse& sy, float dt, std::uint64_t hist, bool pre,
                                         std::int64_t time, std::int64_t age) {
	std::int64_t const prefix = 63 + (pre ? 1 : 0) - time + age;
	// the reference shifts by `prefix` unguarded (UB at 64, only reachable when a source delivers in
	// two consecutive steps); the intended value is taken here: no history bits left
	std::uint64_t const mask = prefix >= 64 ? 0 : (prefix <= 0 ? ~std::uint64_t(0) : ~std::uint64_t(0) >> prefix);
	if (pre)
		f.update(sy, dt, true, (hist & (std::uint64_t(1) << (time - age))) != 0);
	hist &= mask;
	std::int64_t p = prefix;
	while (hist) {
		int const lz = __clzll(static_cast<long long>(hist));
		f.skip(sy, dt, static_cast<Int>(lz - p));
		f.update(sy, dt, false, true);
		hist ^= std::uint64_t(1) << (63 - lz);
		p = lz + 1;
	}
	f.skip(sy, dt, static_cast<Int>(64 - p));
}

// Deliver = true walks the spikes of the delivered step (catch the row's synapses up, count one event per edge for its
// target), several warps to a source; Deliver = false (the 64-step flush, snn.cpp:17-19) walks every source, one warp
// each, and only catches up.
template <class Syn, bool Deliver>
__global__ void __launch_bounds__(256) stateful_visit_kernel(stateful_args a) {
	using S              = typename Syn::synapse;
	int const lane       = threadIdx.x & 31;
	std::int64_t const w = (static_cast<std::int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
	std::int64_t const W = (static_cast<std::int64_t>(gridDim.x) * blockDim.x) >> 5;
	Syn const f          = *static_cast<Syn const*>(a.functor);
	if (Deliver && blockIdx.x == 0 && threadIdx.x == 0)
		*a.evt_cursor = 0;
	unsigned long long ev = 0, sp = 0;
	int const nseg = Deliver ? a.world : 1;
	for (int r = 0; r < nseg; r++) {
		std::int64_t const n = Deliver ? a.ring_cnt[r] : a.n_src;
		if (n == 0)
			continue;
		// Deliver: the step's spiking sources are few (tens to hundreds) and their rows long, so G warps share a source,
		// each taking every G-th group of 32 edges; the source's age is then written by the reserve kernel, after every warp
		// has read it.  The flush walks every source with one warp each and writes the age itself.
		std::int64_t const G      = Deliver ? (W / n > 1 ? W / n : 1) : 1;
		std::int64_t const groups = W / G;
		std::int64_t const g      = w % G;
		for (std::int64_t j = w / G; j < n; j += groups) {
			std::int64_t const src = Deliver ? a.ring_ids[a.seg_lo[r] + j] : j;
			std::int64_t const beg = a.offsets[src], end = a.offsets[src + 1];
			bool pre = false, outdated = false;
			std::int64_t age = a.time + 1;
			if constexpr (PlasticSynapse<Syn>) {
				std::uint64_t const ag = a.ages[src];
				pre                    = (ag >> 63) != 0;
				age                    = static_cast<std::int64_t>(ag & ~(std::uint64_t(1) << 63));
				outdated               = a.time >= age;
			}
			for (std::int64_t e = beg + g * 32 + lane; e < end; e += 32 * G) {
				std::int32_t const dst = a.neighbors[e];
				if constexpr (PlasticSynapse<Syn>) {
					if (outdated) {
						S sy = load_soa<S>(a.syn, a.syn_stride, e);
						catch_up(f, sy, a.dt, a.dst_history[dst], pre, a.time, age);
						store_soa<S>(a.syn, a.syn_stride, e, sy);
					}
				}
				if constexpr (Deliver)
					atomicAdd(a.evt_cnt + dst, 1u);
			}
			if constexpr (PlasticSynapse<Syn> && !Deliver) {
				__syncwarp();
				if (lane == 0)
					a.ages[src] = static_cast<std::uint64_t>(a.time + 1);
			}
			if (g == 0) {
				ev += static_cast<unsigned long long>(end - beg);
				sp++;
			}
		}
	}
	if (Deliver && lane == 0 && sp) {
		atomicAdd(a.stats + 0, ev);
		atomicAdd(a.stats + 1, sp);
	}
}

// every target with events reserves its part of the event list
__global__ void __launch_bounds__(256) stateful_reserve_kernel(stateful_args a) {
	std::int64_t const i = static_cast<std::int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (a.ages) { // plastic: the sources delivered in this step were brought up to it (synapse_population.h:137-138)
		std::int64_t const T = static_cast<std::int64_t>(gridDim.x) * blockDim.x;
		for (int r = 0; r < a.world; r++)
			for (std::int64_t j = i; j < a.ring_cnt[r]; j += T)
				a.ages[a.ring_ids[a.seg_lo[r] + j]] = static_cast<std::uint64_t>(a.time + 1) | (std::uint64_t(1) << 63);
	}
	if (i >= a.n_dst)
		return;
	unsigned const c = a.evt_cnt[i];
	a.evt_fill[i]    = 0;
	if (c) {
		unsigned long long const at = atomicAdd(a.evt_cursor, static_cast<unsigned long long>(c));
		if (at + c > static_cast<unsigned long long>(a.evt_cap)) {
			atomicOr(a.error, 32);
			a.evt_cnt[i] = 0; // dropped (reported as an error by the next synchronising call)
		} else
			a.evt_off[i] = static_cast<std::uint32_t>(at);
	}
}

__global__ void __launch_bounds__(256) stateful_fill_kernel(stateful_args a) {
	int const lane       = threadIdx.x & 31;
	std::int64_t const w = (static_cast<std::int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
	std::int64_t const W = (static_cast<std::int64_t>(gridDim.x) * blockDim.x) >> 5;
	for (int r = 0; r < a.world; r++) {
		std::int64_t const n = a.ring_cnt[r];
		if (n == 0)
			continue;
		std::int64_t const G      = W / n > 1 ? W / n : 1; // warps per source, as in the visit
		std::int64_t const groups = W / G;
		std::int64_t const g      = w % G;
		for (std::int64_t j = w / G; j < n; j += groups) {
			std::int64_t const src = a.ring_ids[a.seg_lo[r] + j];
			std::int64_t const beg = a.offsets[src], end = a.offsets[src + 1];
			for (std::int64_t e = beg + g * 32 + lane; e < end; e += 32 * G) {
				std::int32_t const dst = a.neighbors[e];
				if (a.evt_cnt[dst]) {
					unsigned const k                   = atomicAdd(a.evt_fill + dst, 1u);
					a.evt_list[a.evt_off[dst] + k] = static_cast<std::int32_t>(e);
				}
			}
		}
	}
}

// ---- stateful neurons: one thread per neuron, whole window --------------------------------------
// NIN = number of incoming connections the loops are unrolled for; Generic = false promises that
// a.n_in == NIN, that every incoming connection is a stateless one fed by the tiled delivery kernel
// (counters never need zeroing) — the common case, with nothing but the loads, the calls and the
// model's own arithmetic left in the loop.  Generic = true handles everything else (stateful
// synapses' event lists, the atomic delivery mode, up to kMaxIncoming connections).
template <class FSyn, int C>
struct fused_synapses {
	FSyn v[C];
};
template <int C>
struct fused_synapses<void, C> {};

//
// FSyn != void: every incoming connection carries the same stateless synapse type, whose deliver()
// is inlined (the type-erased apply is an indirect call that moves the neuron through local memory).
template <class Neur, int NIN, bool Generic, class FSyn = void>
__global__ void __launch_bounds__(128, 8) update_stateful_kernel(update_args a) {
	using N                  = typename Neur::neuron;
	std::int64_t const i     = static_cast<std::int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
	bool const active        = i < a.n_local;
	std::int64_t const ii    = active ? i : 0;
	unsigned const lane_lt   = (1u << (threadIdx.x & 31)) - 1;
	Neur const neur          = *static_cast<Neur const*>(a.functor);
	N n                      = load_soa<N>(a.state, a.stride, ii);
	std::uint64_t hist       = (a.history && active) ? a.history[ii] : 0;
	std::int32_t const my_id = static_cast<std::int32_t>(a.lo + i);
	// a neuron that draws: its draws of step t sit at (population offset + global index * draws) of the step's stream
	// (neuron_population.h:116-124 walks the neurons in order with the step's one engine): the chunk's jump polynomial
	// gives the engine at the chunk's first neuron, this neuron's own position is (index in chunk) * draws steps on
	constexpr int draws = rng_draws_v<Neur>;
	u128 my_poly{0, 0};
	if constexpr (draws > 0)
		my_poly = a.jump_poly[ii / kRngChunk];
	constexpr int C = NIN > 0 ? NIN : 1;

	// event counters of the step, one per incoming connection; the next step's are fetched while
	// this step's are applied (they were all written before this window began).  All incoming
	// connections share one counter ring length; slots advance with the step.
	std::uint32_t const* cp[C];
	unsigned kk[C];
#pragma unroll
	for (int c = 0; c < C; c++) {
		cp[c] = nullptr;
		kk[c] = 0;
		if (c < NIN && active && (!Generic || (c < a.n_in && !a.in[c].evt_cnt))) {
			cp[c] = a.in[c].counts + ii;
			kk[c] = cp[c][static_cast<std::int64_t>(a.cslot0) * a.in[c].cstride];
		}
	}
	int cslot = a.cslot0, rslot = a.rslot0;
	// the connections' synapse objects (their parameters), read once
	fused_synapses<FSyn, C> fsyn;
	if constexpr (!std::is_void_v<FSyn>) {
#pragma unroll
		for (int c = 0; c < C; c++)
			if (c < NIN)
				fsyn.v[c] = *static_cast<FSyn const*>(a.in[c].functor);
	}
	// Two incoming connections in a row with the SAME synapse object (samples/brunel: P->E and E->E both carry w_exc): the k1
	// deliveries of one followed by the k2 of the other are k1 + k2 calls of the same function — the same float operations in
	// the same order, one loop instead of two.
	bool same_as_next[C] = {};
	if constexpr (!std::is_void_v<FSyn> && !Generic) {
#pragma unroll
		for (int c = 0; c + 1 < C; c++)
			if (c + 1 < NIN) {
				bool eq = !a.in[c].zero_after_read && !a.in[c + 1].zero_after_read;
				unsigned char const* x = reinterpret_cast<unsigned char const*>(&fsyn.v[c]);
				unsigned char const* y = reinterpret_cast<unsigned char const*>(&fsyn.v[c + 1]);
				for (unsigned k = 0; k < sizeof(FSyn); k++)
					eq = eq && x[k] == y[k];
				same_as_next[c] = eq;
			}
	}

	for (int s = 0; s < a.nsteps; s++) {
		int const cnext = cslot + 1 == a.cring ? 0 : cslot + 1;
		unsigned kn[C];
#pragma unroll
		for (int c = 0; c < C; c++) {
			kn[c] = 0;
			if (c < NIN && s + 1 < a.nsteps && cp[c])
				kn[c] = cp[c][static_cast<std::int64_t>(cnext) * a.in[c].cstride];
		}
		// fold in the events whose delivery the reference ran at the end of step t-1
		if constexpr (!std::is_void_v<FSyn> && !Generic) {
#pragma unroll
			for (int c = 0; c + 1 < C; c++)
				if (c + 1 < NIN && same_as_next[c]) {
					kk[c + 1] += kk[c];
					kk[c] = 0;
				}
		}
#pragma unroll
		for (int c = 0; c < C; c++) {
			if (c < NIN && kk[c]) {
				incoming const& in = a.in[c];
				if (in.zero_after_read) // the producer adds to these counters (atomic delivery, or tiled delivery in rounds)
					in.counts[cslot * in.cstride + ii] = 0;
				if constexpr (std::is_void_v<FSyn>)
					in.apply(in.functor, &n, kk[c]);
				else
					deliver_k(fsyn.v[c], n, kk[c]);
			}
			if constexpr (Generic) {
				if (c < a.n_in && a.in[c].evt_cnt && active) { // stateful synapses: this step's event list (window = 1 step)
					incoming const& in = a.in[c];
					unsigned const ne  = in.evt_cnt[ii];
					if (ne) {
						in.evt_cnt[ii] = 0;
						in.apply_events(in.functor, &n, in.syn, in.syn_stride, in.evt_list + in.evt_off[ii], ne, &in.from);
					}
				}
			}
			kk[c] = kn[c];
		}
		bool spiked;
		if constexpr (draws > 0) {
			u128 const* nib = a.rng.nib + static_cast<std::int64_t>(s) * 32 * 16;
			UInt s0 = 0, s1 = 0;
#pragma unroll 8
			for (int g = 0; g < 32; g++) {
				unsigned const v   = static_cast<unsigned>(((g < 16 ? my_poly.lo : my_poly.hi) >> (4 * (g & 15))) & 15);
				ulonglong2 const e = __ldg(reinterpret_cast<ulonglong2 const*>(nib + g * 16 + v));
				s0 ^= e.x;
				s1 ^= e.y;
			}
			counting_rng rng;
			rng.g = util::xoroshiro64_128p(s0, s1);
			for (int k = static_cast<int>(ii % kRngChunk) * draws; k > 0; k--)
				rng.g.advance();
			spiked = active && neur.update(n, a.dt[s], rng);
			if (rng.used > draws)
				atomicOr(a.error, 64);
		} else {
			null_rng rng;
			spiked = active && neur.update(n, a.dt[s], rng);
		}
		hist = (hist << 1) | (spiked ? 1u : 0u);

		unsigned const m = __ballot_sync(0xffffffffu, spiked);
		if (m) {
			std::int64_t const slot = rslot;
			unsigned base           = 0;
			if ((threadIdx.x & 31) == 0)
				base = atomicAdd(&a.ring_cnt[slot * a.world + a.rank], __popc(m));
			base = __shfl_sync(0xffffffffu, base, 0);
			if (spiked) {
				std::int64_t const at = slot * a.ring_cap + a.lo + base + __popc(m & lane_lt);
				for (int r = 0; r < a.world; r++)
					a.ring_ids[r][at] = my_id;
			}
		}
		cslot = cnext;
		rslot = rslot + 1 == a.ring ? 0 : rslot + 1;
	}
	if (active) {
		store_soa<N>(a.state, a.stride, i, n);
		if (a.history)
			a.history[i] = hist;
	}
}

// ---- stateless neurons: one thread per (step, 32 consecutive neurons) ----------------------------
// A flat grid, the step's lookup tables read through L1, one returning atomic per spike.  A variant with one step per grid
// row, the tables staged in shared memory and ONE atomic per warp for its spikes ran 3 % faster alone (17.1 against 17.6 us
// per window of the bench's 353,555 Poisson neurons) and made the update phase 9 us SLOWER (81.8 against 72.9 us, A/B on one
// box): this kernel shares the SMs with the stateful populations' update kernels on other streams, and what counts is how
// its blocks pack beside theirs, not its own duration.  Removed.
template <class Neur>
__global__ void __launch_bounds__(256) update_stateless_kernel(update_args a) {
	constexpr int draws        = rng_draws_v<Neur>;
	std::int64_t const chunks  = (a.n_local + kRngChunk - 1) / kRngChunk;
	std::int64_t const item    = static_cast<std::int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (item >= chunks * a.nsteps)
		return;
	int const s              = static_cast<int>(item / chunks);
	std::int64_t const chunk = item % chunks;
	std::int64_t const t     = a.t0 + s;
	Neur const neur          = *static_cast<Neur const*>(a.functor);
	float const dt           = a.dt[s];

	counting_rng rng;
	if constexpr (draws > 0) {
		// state at this chunk's offset in step t's stream: XOR of the basis states selected by
		// the chunk's jump polynomial, 4 coefficients per lookup
		u128 const poly = a.jump_poly[chunk];
		u128 const* nib = a.rng.nib + static_cast<std::int64_t>(s) * 32 * 16;
		UInt s0 = 0, s1 = 0;
#pragma unroll 8
		for (int g = 0; g < 32; g++) {
			unsigned const v = static_cast<unsigned>(((g < 16 ? poly.lo : poly.hi) >> (4 * (g & 15))) & 15);
			ulonglong2 const e = __ldg(reinterpret_cast<ulonglong2 const*>(nib + g * 16 + v));
			s0 ^= e.x;
			s1 ^= e.y;
		}
		rng.g = util::xoroshiro64_128p(s0, s1);
	}

	std::int64_t const first = chunk * kRngChunk;
	std::int64_t const slot  = t % a.ring;
	for (int q = 0; q < kRngChunk; q++) {
		std::int64_t const i = first + q;
		if (i >= a.n_local)
			break;
		rng.used = 0;
		bool spiked;
		if constexpr (draws > 0)
			spiked = neur.update(dt, rng);
		else {
			null_rng none;
			spiked = neur.update(dt, none);
		}
		if constexpr (draws > 0) {
			// fewer draws than declared: skip the rest so the next neuron starts at its own position.  More: every neuron
			// behind this one in the chunk would read the wrong part of the stream — reported, not papered over
			if (rng.used > draws)
				atomicOr(a.error, 64);
			for (; rng.used < draws; rng.used++)
				rng.g.advance();
		}
		if (spiked) {
			unsigned const pos    = atomicAdd(&a.ring_cnt[slot * a.world + a.rank], 1u);
			std::int64_t const at = slot * a.ring_cap + a.lo + pos;
			for (int r = 0; r < a.world; r++)
				a.ring_ids[r][at] = static_cast<std::int32_t>(a.lo + i);
		}
	}
}

// ---- state export / import (neuron_population::get_neurons, neuron_population.h:142-145) ---------
// Export folds the pending events of step t_next into a COPY of the state: the reference applies
// them at the end of the step that was just run, this backend at the start of the next one.
template <class Neur>
__global__ void __launch_bounds__(256) export_kernel(export_args a) {
	using N              = typename Neur::neuron;
	std::int64_t const i = static_cast<std::int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (i >= a.n_local)
		return;
	N n = load_soa<N>(a.state, a.stride, i);
	for (int c = 0; c < a.n_in; c++) {
		incoming const& in = a.in[c];
		if (in.evt_cnt) {
			unsigned const ne = in.evt_cnt[i];
			if (ne)
				in.apply_events(in.functor, &n, in.syn, in.syn_stride, in.evt_list + in.evt_off[i], ne, &in.from);
			continue;
		}
		unsigned const k = in.counts[(a.t_next % in.ring) * in.cstride + i];
		if (k)
			in.apply(in.functor, &n, k);
	}
	static_cast<N*>(a.out_aos)[i] = n;
}

template <class Neur>
__global__ void __launch_bounds__(256) import_kernel(import_args a) {
	using N              = typename Neur::neuron;
	std::int64_t const i = static_cast<std::int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (i >= a.n_local)
		return;
	store_soa<N>(a.state, a.stride, i, static_cast<N const*>(a.in_aos)[i]);
	for (int c = 0; c < a.n_in; c++) { // the imported state already holds the events export_kernel folded in
		incoming const& in = a.in[c];
		if (in.evt_cnt)
			in.evt_cnt[i] = 0;
		else
			in.counts[(a.t_next % in.ring) * in.cstride + i] = 0;
	}
}

inline int grid_for(std::int64_t threads, int block = 256) {
	return static_cast<int>((threads + block - 1) / block);
}

// ---- ops tables ----------------------------------------------------------------------------------
template <Neuron Neur>
struct neuron_ops_builder {
	static void init_host(void const* functor, void* out, std::int64_t n, std::uint64_t seed_lo, std::uint64_t seed_hi) {
		if constexpr (StatefulNeuron<Neur>) {
			using N = typename Neur::neuron;
			N* v    = static_cast<N*>(out);
			for (std::int64_t i = 0; i < n; i++)
				new (v + i) N();
			Neur neur = *static_cast<Neur const*>(functor);
			util::xoroshiro64_128p rng(seed_lo, seed_hi);
			if constexpr (PerNeuronInit<Neur>) {
				for (std::int64_t i = 0; i < n; i++)
					neur.init(v[i], i, rng);
			} else if constexpr (PerPopulationInit<Neur>) {
				neur.init(std::span<N>(v, static_cast<std::size_t>(n)), rng);
			}
		}
		(void)functor, (void)out, (void)n, (void)seed_lo, (void)seed_hi;
	}

	static int launch_update(update_args const* a) {
		auto stream = static_cast<cudaStream_t>(a->stream);
		if constexpr (StatefulNeuron<Neur>) {
			if (a->n_local > 0) {
				bool fast = a->n_in <= 4;
				for (int c = 0; c < a->n_in; c++)
					fast = fast && !a->in[c].evt_cnt;
				int const grid = grid_for(a->n_local, 128);
				if (!fast)
					update_stateful_kernel<Neur, kMaxIncoming, true><<<grid, 128, 0, stream>>>(*a);
				else
					switch (a->n_in) {
					case 0: update_stateful_kernel<Neur, 0, false><<<grid, 128, 0, stream>>>(*a); break;
					case 1: update_stateful_kernel<Neur, 1, false><<<grid, 128, 0, stream>>>(*a); break;
					case 2: update_stateful_kernel<Neur, 2, false><<<grid, 128, 0, stream>>>(*a); break;
					case 3: update_stateful_kernel<Neur, 3, false><<<grid, 128, 0, stream>>>(*a); break;
					default: update_stateful_kernel<Neur, 4, false><<<grid, 128, 0, stream>>>(*a); break;
					}
			}
		} else {
			std::int64_t const chunks = (a->n_local + kRngChunk - 1) / kRngChunk;
			if (chunks > 0 && a->nsteps > 0)
				update_stateless_kernel<Neur><<<grid_for(chunks * a->nsteps), 256, 0, stream>>>(*a);
		}
		return static_cast<int>(cudaGetLastError());
	}

	static int launch_export(export_args const* a) {
		if constexpr (StatefulNeuron<Neur>) {
			if (a->n_local > 0)
				export_kernel<Neur><<<grid_for(a->n_local), 256, 0, static_cast<cudaStream_t>(a->stream)>>>(*a);
		}
		return static_cast<int>(cudaGetLastError());
	}

	static int launch_import(import_args const* a) {
		if constexpr (StatefulNeuron<Neur>) {
			if (a->n_local > 0)
				import_kernel<Neur><<<grid_for(a->n_local), 256, 0, static_cast<cudaStream_t>(a->stream)>>>(*a);
		}
		return static_cast<int>(cudaGetLastError());
	}

	static constexpr std::uint32_t neuron_bytes() {
		if constexpr (StatefulNeuron<Neur>)
			return sizeof(typename Neur::neuron);
		else
			return 0;
	}
};

template <Neuron Neur>
spice_neuron_ops const* neuron_ops(char const* name = "user") {
	static_assert(!PerPopulationUpdate<Neur>, "per-population update() neurons are host-fed; not on the GPU path yet");
	static_assert(std::is_trivially_copyable_v<Neur>, "neuron functors are copied to the device byte-wise");
	using B = neuron_ops_builder<Neur>;
	static spice_neuron_ops const ops{1,
	                                  name,
	                                  B::neuron_bytes(),
	                                  static_cast<std::uint32_t>(sizeof(Neur)),
	                                  static_cast<std::uint32_t>(rng_draws_v<Neur>),
	                                  StatefulNeuron<Neur> ? 1u : 0u,
	                                  &B::init_host,
	                                  &B::launch_update,
	                                  &B::launch_export,
	                                  &B::launch_import};
	return &ops;
}

// update kernel of DstNeur with Syn::deliver inlined, for populations whose incoming connections all
// carry this synapse type (whatever their sources): one function, hence one address, per (Syn, DstNeur)
template <class Syn, StatefulNeuron DstNeur>
int launch_update_fused(update_args const* a) {
	if constexpr (!StatefulSynapse<Syn> && DeliverTo<Syn, DstNeur> && rng_draws_v<DstNeur> == 0) {
		auto stream    = static_cast<cudaStream_t>(a->stream);
		int const grid = grid_for(a->n_local, 128);
		if (a->n_local > 0)
			switch (a->n_in) {
			case 1: update_stateful_kernel<DstNeur, 1, false, Syn><<<grid, 128, 0, stream>>>(*a); break;
			case 2: update_stateful_kernel<DstNeur, 2, false, Syn><<<grid, 128, 0, stream>>>(*a); break;
			case 3: update_stateful_kernel<DstNeur, 3, false, Syn><<<grid, 128, 0, stream>>>(*a); break;
			default: update_stateful_kernel<DstNeur, 4, false, Syn><<<grid, 128, 0, stream>>>(*a); break;
			}
		return static_cast<int>(cudaGetLastError());
	} else {
		(void)a;
		return -1;
	}
}

// A STATELESS synapse whose deliver() reads the source neuron (concepts.h:76-99, synapse_population.h:125-131): its events
// differ from source to source, so they cannot be counted; they take the event-list path of the stateful synapses
// (per target, in the reference's (source, row) order) with one unused word of state per edge.
template <class Syn, class SrcNeur, class DstNeur>
struct carried_from_to : Syn {
	struct synapse {
		std::int32_t unused = 0;
	};
	SPICE_HD void deliver(synapse const&, typename SrcNeur::neuron const& from, typename DstNeur::neuron& to) const {
		Syn::deliver(from, to);
	}
};

template <class Syn, Neuron SrcNeur, StatefulNeuron DstNeur>
requires Synapse<Syn, SrcNeur, DstNeur>
spice_synapse_ops const* synapse_ops(char const* name = "user") {
	static_assert(std::is_trivially_copyable_v<Syn>, "synapse functors are copied to the device byte-wise");
	if constexpr (!StatefulSynapse<Syn> && detail::deliver_from_to_v<Syn, SrcNeur, DstNeur>) {
		using carried = carried_from_to<Syn, SrcNeur, DstNeur>;
		static_assert(sizeof(carried) == sizeof(Syn) && Synapse<carried, SrcNeur, DstNeur>);
		return synapse_ops<carried, SrcNeur, DstNeur>(name);
	} else {
	struct B {
		static int get_apply(apply_fn* out) {
			return static_cast<int>(cudaMemcpyFromSymbol(out, apply_ptr<Syn, DstNeur>, sizeof(apply_fn)));
		}
		static int get_apply_events(apply_events_fn* out) {
			return static_cast<int>(cudaMemcpyFromSymbol(out, apply_events_ptr<Syn, SrcNeur, DstNeur>, sizeof(apply_events_fn)));
		}
		static void init_host(void const* functor, void* out, std::int64_t const* offsets, std::int32_t const* neighbors,
		                      std::int64_t n_src, std::uint64_t seed_lo, std::uint64_t seed_hi) {
			if constexpr (StatefulSynapse<Syn>) {
				using S = typename Syn::synapse;
				S* v    = static_cast<S*>(out);
				for (std::int64_t e = 0; e < offsets[n_src]; e++)
					new (v + e) S();
				if constexpr (PerSynapseInit<Syn>) { // synapse_population.h:34-41
					Syn const f = *static_cast<Syn const*>(functor);
					util::xoroshiro64_128p rng(seed_lo, seed_hi);
					for (std::int64_t src = 0; src < n_src; src++)
						for (std::int64_t e = offsets[src]; e < offsets[src + 1]; e++)
							f.init(v[e], src, neighbors[e], rng);
				}
			}
			(void)functor, (void)out, (void)offsets, (void)neighbors, (void)n_src, (void)seed_lo, (void)seed_hi;
		}
		static int launch_stateful(stateful_args const* a) {
			if constexpr (StatefulSynapse<Syn>) {
				auto stream = static_cast<cudaStream_t>(a->stream);
				switch (a->phase) {
				case 0: stateful_visit_kernel<Syn, true><<<148 * 4, 256, 0, stream>>>(*a); break;
				case 1:
					if (a->n_dst > 0 || a->ages)
						stateful_reserve_kernel<<<grid_for(a->n_dst > 0 ? a->n_dst : 1), 256, 0, stream>>>(*a);
					break;
				case 2: stateful_fill_kernel<<<148 * 4, 256, 0, stream>>>(*a); break;
				case 3:
					if constexpr (PlasticSynapse<Syn>)
						stateful_visit_kernel<Syn, false><<<148 * 8, 256, 0, stream>>>(*a);
					break;
				}
			}
			(void)a;
			return static_cast<int>(cudaGetLastError());
		}
	};
	constexpr std::uint32_t syn_bytes = [] {
		if constexpr (StatefulSynapse<Syn>)
			return static_cast<std::uint32_t>(sizeof(typename Syn::synapse));
		else
			return 0u;
	}();
	static spice_synapse_ops const ops{1,
	                                   name,
	                                   syn_bytes,
	                                   static_cast<std::uint32_t>(sizeof(Syn)),
	                                   static_cast<std::uint32_t>(sizeof(typename DstNeur::neuron)),
	                                   PlasticSynapse<Syn> ? 1u : 0u,
	                                   DeliverTo<Syn, DstNeur> ? 0u : 1u,
	                                   &B::get_apply,
	                                   PerSynapseInit<Syn> ? 1u : 0u,
	                                   &B::init_host,
	                                   &B::get_apply_events,
	                                   &B::launch_stateful,
	                                   (!StatefulSynapse<Syn> && DeliverTo<Syn, DstNeur>) ? &launch_update_fused<Syn, DstNeur> : nullptr};
	return &ops;
	}
}
}
