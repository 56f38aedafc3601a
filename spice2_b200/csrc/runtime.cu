// Model-agnostic runtime of the B200 backend and the C ABI (include/spice_b200.h).
//
// It stands where the reference has `class snn` + the type-erased NeuronPopulation /
// SynapsePopulation interfaces (spice/include/spice/snn.h:16-75, spice/src/snn.cpp:7-28,
// spice/include/spice/detail/{neuron_population,synapse_population}.h): it owns the populations,
// the connections and the seed bookkeeping, and sequences update -> (exchange) -> delivery.
// Unlike the reference it does so per WINDOW of up to min-delay steps: a spike emitted inside a
// window cannot be consumed inside it (snn.cpp:21-25: emitted at t, first visible at t+delay),
// so one launch per population updates all steps of the window and one launch per connection
// delivers all of the window's spikes into future per-step event counters.
//
// The kernels that touch user functors are instantiated elsewhere (spice/detail/model_ops.cuh)
// and reach this file through the ops tables; the kernels here are functor-free: spike
// delivery into integer counters, window prologue (per-step RNG jump tables), peer publication /
// wait for the multi-GPU spike exchange over NVLink, raster packing.
#include <cuda_runtime.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "deliver.h"
#include "generator.h"
#include "spice/detail/abi.h"
#include "spice/util/numeric.h"
#include "spice/util/random.h"
#include "spice_b200.h"

using namespace spice;
using namespace spice::detail;

namespace {
thread_local std::string g_create_error;

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------

// Window prologue: (1) zero this rank's spike counters for the window's steps, (2) for every step
// build the 4-bit jump tables of that step's xoroshiro stream:
// nib[s][g][v] = XOR_{b in v} T^(4g+b) seed_s.  One block per step.
struct prologue_args {
	u128 seed[kMaxWindow]; // seed of each step's stream (snn.cpp:12: rng(_seed++))
	u128* nib;             // [nsteps][32][16] or null when no population draws
	std::uint32_t* const* ring_cnt; // device array: per population, counters [ring][world]
	int npops;
	int ring, world, rank;
	long long t0;
	int nsteps;
	unsigned* work; // unit counter of the tiled delivery kernel
};

__global__ void __launch_bounds__(512) window_prologue(prologue_args a) {
	int const s = blockIdx.x;
	if (s == 0 && threadIdx.x == 0 && a.work)
		*a.work = 0;
	for (int p = threadIdx.x; p < a.npops; p += blockDim.x)
		a.ring_cnt[p][((a.t0 + s) % a.ring) * a.world + a.rank] = 0;
	if (!a.nib)
		return;
	__shared__ ulonglong2 basis[128];
	if (threadIdx.x == 0) {
		unsigned long long s0 = a.seed[s].lo, s1 = a.seed[s].hi;
		for (int i = 0; i < 128; i++) {
			basis[i]                   = make_ulonglong2(s0, s1);
			unsigned long long const t = s0 ^ s1;
			s0                         = ((s0 << 24) | (s0 >> 40)) ^ t ^ (t << 16);
			s1                         = (t << 37) | (t >> 27);
		}
	}
	__syncthreads();
	int const g = threadIdx.x >> 4, v = threadIdx.x & 15;
	unsigned long long x = 0, y = 0;
	for (int b = 0; b < 4; b++)
		if (v & (1 << b)) {
			x ^= basis[4 * g + b].x;
			y ^= basis[4 * g + b].y;
		}
	a.nib[(static_cast<long long>(s) * 32 + g) * 16 + v] = u128{x, y};
}

// Spike delivery for connections whose events are integer counts (all stateless synapses):
// the reference's hot loop B (synapse_population.h:88,99,118-133) — for src in spikes, for dst
// in row(src): deliver — with `deliver` deferred: one warp per spiking source streams that
// source's CSR row and bumps counts[slot(step + delay)][dst].
struct deliver_args {
	std::int32_t const* ring_ids;  // source population spike ring (this rank's copy)
	std::uint32_t const* ring_cnt; // [ring][world]
	long long ring_cap;
	int ring, world;
	long long seg_lo[kMaxWorld]; // first neuron of each rank's range in the source population
	long long const* offsets;    // CSR (rows = all sources, local columns)
	std::int32_t const* neighbors;
	std::uint32_t* counts;       // [cring][cstride]
	long long cstride;
	int cring;
	long long delay;
	long long t0;
	int nsteps;
	unsigned long long* stats; // [0] events, [1] spikes
};

__global__ void __launch_bounds__(256) deliver_counts(deliver_args a) {
	__shared__ unsigned prefix[kMaxWindow * kMaxWorld + 1];
	int const nseg = a.nsteps * a.world;
	if (threadIdx.x == 0) {
		unsigned run = 0;
		for (int i = 0; i < nseg; i++) {
			prefix[i]   = run;
			int const s = i / a.world, r = i % a.world;
			run += a.ring_cnt[((a.t0 + s) % a.ring) * a.world + r];
		}
		prefix[nseg] = run;
	}
	__syncthreads();
	unsigned const total = prefix[nseg];
	int const lane       = threadIdx.x & 31;
	unsigned const warps = gridDim.x * (blockDim.x >> 5);
	unsigned long long ev = 0, sp = 0;
	for (unsigned item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); item < total; item += warps) {
		// which (step, rank) segment does this spike belong to
		int lo = 0, hi = nseg - 1;
		while (lo < hi) {
			int const mid = (lo + hi + 1) >> 1;
			if (prefix[mid] <= item)
				lo = mid;
			else
				hi = mid - 1;
		}
		int const s = lo / a.world, r = lo % a.world;
		long long const t   = a.t0 + s;
		std::int32_t const src = a.ring_ids[(t % a.ring) * a.ring_cap + a.seg_lo[r] + (item - prefix[lo])];
		long long const beg = a.offsets[src], end = a.offsets[src + 1];
		std::uint32_t* cnt  = a.counts + ((t + a.delay) % a.cring) * a.cstride;
		long long e = beg + lane;
		for (; e + 96 < end; e += 128) {
			std::int32_t const d0 = a.neighbors[e], d1 = a.neighbors[e + 32], d2 = a.neighbors[e + 64],
			                   d3 = a.neighbors[e + 96];
			atomicAdd(cnt + d0, 1u);
			atomicAdd(cnt + d1, 1u);
			atomicAdd(cnt + d2, 1u);
			atomicAdd(cnt + d3, 1u);
		}
		for (; e < end; e += 32)
			atomicAdd(cnt + a.neighbors[e], 1u);
		ev += static_cast<unsigned long long>(end - beg);
		sp++;
	}
	if (lane == 0 && sp) {
		atomicAdd(a.stats + 0, ev);
		atomicAdd(a.stats + 1, sp);
	}
}

// Multi-GPU: after this rank's update kernels of window `seq` finished, copy its spike counters
// into every peer's counter table and raise this rank's flag there.  The spike ids themselves
// were stored into the peers' rings by the update kernels (NVLink peer stores).
struct publish_args {
	std::uint32_t* const* local_cnt;               // [npops] this rank's tables
	std::uint32_t* const* peer_cnt;                // [world][npops] device array of peers' tables
	unsigned long long* peer_flags[kMaxWorld];     // peers' flag arrays [world]
	int npops, ring, world, rank;
	long long t0;
	int nsteps;
	unsigned long long seq;
};

__global__ void __launch_bounds__(256) publish_window(publish_args a) {
	int const n = a.npops * a.nsteps * a.world;
	for (int i = threadIdx.x; i < n; i += blockDim.x) {
		int const q = i % a.world, s = (i / a.world) % a.nsteps, p = i / (a.world * a.nsteps);
		if (q == a.rank)
			continue;
		long long const at       = ((a.t0 + s) % a.ring) * a.world + a.rank;
		a.peer_cnt[q * a.npops + p][at] = a.local_cnt[p][at];
	}
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x < a.world) {
		volatile unsigned long long* f = a.peer_flags[threadIdx.x] + a.rank;
		*f                             = a.seq;
	}
}

struct wait_args {
	unsigned long long const* flags; // this rank's flag array [world]
	int world;
	unsigned long long seq;
	int* error;
};

__device__ __forceinline__ unsigned long long wall_ns() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}

__global__ void wait_window(wait_args a) {
	if (static_cast<int>(threadIdx.x) >= a.world)
		return;
	volatile unsigned long long const* f = a.flags + threadIdx.x;
	unsigned long long const start       = wall_ns();
	while (*f < a.seq) {
		if (wall_ns() - start > 10000000000ull) { // 10 s of wall time (%globaltimer: independent of the SM clock): a peer died; do not hang the GPU
			atomicOr(a.error, 1);
			break;
		}
		__nanosleep(200);
	}
	__threadfence_system();
}

// Spike sink (the per-step sink of the samples, samples/matplot.cpp:96-134, as a
// batched ring drained by the host without a device-wide sync): one block per (step, population)
// of the window sorts the step's spike list ascending — the order neuron_population::spikes()
// has in the reference — and appends it to a ring in page-locked, device-mapped HOST memory.
// Sorting is a bitmap over the population's id range in shared memory (ids are unique), read
// back with popc prefix sums.  The lists of a window sit in (step, population) order, so a
// readout of n steps is at most two memcpy()s on the host.
constexpr int kSinkThreads = 1024;
struct sink_args {
	std::int32_t const* const* ring_ids; // [npops]
	std::uint32_t const* const* ring_cnt;
	long long const* ring_cap;           // [npops]
	long long const* seg_lo;             // [npops][world]
	long long const* pop_size;           // [npops]
	int npops, ring, world;
	long long t0;
	int nsteps;
	int parity;                          // window index & 1: cursor[parity] is this window's base
	long long step_index0;               // index of the window's first step since the sink was enabled
	unsigned long long* cursor;          // [2] ids appended before this / the next window
	std::int32_t* stage;                 // device ring, same positions as the host ring
	std::int32_t* h_ids;                 // host ring (mapped)
	long long cap;                       // ids in either ring
	long long* h_cnt;                    // [steps_cap][npops] (mapped)
	unsigned long long* h_off;           // [steps_cap][npops] (mapped) monotonic position of the list
	long long steps_cap;
	unsigned long long consumed_steps, consumed_ids; // what the host had taken when the window was enqueued (a lower bound)
	int* h_error;                        // mapped
};

__global__ void __launch_bounds__(kSinkThreads) sink_pack(sink_args a) {
	extern __shared__ unsigned sink_bm[];
	__shared__ unsigned long long before_s, base_s;
	__shared__ unsigned total_s, warp_sum[32];
	__shared__ int ok_s;
	int const tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	int const s = blockIdx.x / a.npops, p = blockIdx.x % a.npops;
	long long const si   = a.step_index0 + s;
	long long const slot = (a.t0 + s) % a.ring;
	if (tid == 0)
		before_s = 0, total_s = 0;
	__syncthreads();
	// where this list starts: after the lists of the window's earlier (step, population) pairs
	for (int b = tid; b <= static_cast<int>(blockIdx.x); b += kSinkThreads) {
		int const s2 = b / a.npops, p2 = b % a.npops;
		long long const slot2 = (a.t0 + s2) % a.ring;
		unsigned t = 0;
		for (int r = 0; r < a.world; r++)
			t += a.ring_cnt[p2][slot2 * a.world + r];
		if (b < static_cast<int>(blockIdx.x))
			atomicAdd(&before_s, static_cast<unsigned long long>(t));
		else
			total_s = t;
	}
	__syncthreads();
	if (tid == 0) {
		unsigned long long const base = a.cursor[a.parity] + before_s;
		if (blockIdx.x == gridDim.x - 1)
			a.cursor[a.parity ^ 1] = base + total_s;
		int ok = 1;
		if (static_cast<unsigned long long>(si) - a.consumed_steps >= static_cast<unsigned long long>(a.steps_cap))
			*a.h_error = 4, ok = 0;
		else if (base + total_s - a.consumed_ids > static_cast<unsigned long long>(a.cap))
			*a.h_error = 8, ok = 0;
		if (ok) {
			a.h_cnt[(si % a.steps_cap) * a.npops + p] = total_s;
			a.h_off[(si % a.steps_cap) * a.npops + p] = base;
		}
		base_s = base;
		ok_s   = ok;
	}
	__syncthreads();
	if (!ok_s)
		return;
	unsigned long long const base = base_s;
	unsigned const total          = total_s;
	long long const size          = a.pop_size[p];
	// words per thread: odd, so that a warp's strided bitmap reads fall into different banks
	int const W            = static_cast<int>(min(31ll, ((size + 32 * kSinkThreads - 1) / (32 * kSinkThreads)) | 1));
	int const words        = W * kSinkThreads;
	long long const cbits  = static_cast<long long>(words) * 32;
	unsigned emitted       = 0;
	for (long long c0 = 0; c0 < size && emitted < total; c0 += cbits) {
		for (int i = tid; i < words; i += kSinkThreads)
			sink_bm[i] = 0;
		__syncthreads();
		for (int r = 0; r < a.world; r++) {
			unsigned const c         = a.ring_cnt[p][slot * a.world + r];
			std::int32_t const* from = a.ring_ids[p] + slot * a.ring_cap[p] + a.seg_lo[p * a.world + r];
			for (unsigned j = tid; j < c; j += kSinkThreads) {
				long long const id = from[j] - c0;
				if (id >= 0 && id < cbits)
					atomicOr(&sink_bm[id >> 5], 1u << (id & 31));
			}
		}
		__syncthreads();
		unsigned mine = 0;
		for (int k = 0; k < W; k++)
			mine += __popc(sink_bm[tid * W + k]);
		unsigned incl = mine;
		for (int off = 1; off < 32; off <<= 1) {
			unsigned const o = __shfl_up_sync(0xffffffffu, incl, off);
			if (lane >= off)
				incl += o;
		}
		if (lane == 31)
			warp_sum[warp] = incl;
		__syncthreads();
		if (warp == 0) {
			unsigned v = warp_sum[lane], w = v;
			for (int off = 1; off < 32; off <<= 1) {
				unsigned const o = __shfl_up_sync(0xffffffffu, w, off);
				if (lane >= off)
					w += o;
			}
			warp_sum[lane] = w - v; // exclusive
			if (lane == 31)
				total_s = w; // ids of this chunk
		}
		__syncthreads();
		unsigned long long pos = base + emitted + warp_sum[warp] + (incl - mine);
		for (int k = 0; k < W; k++) {
			unsigned m = sink_bm[tid * W + k];
			while (m) {
				int const bit = __ffs(m) - 1;
				m &= m - 1;
				a.stage[pos % a.cap] = static_cast<std::int32_t>(c0 + (static_cast<long long>(tid) * W + k) * 32 + bit);
				pos++;
			}
		}
		emitted += total_s;
		__syncthreads();
	}
	// coalesced copy of the sorted list into the host ring (128-byte PCIe writes)
	for (unsigned k = tid; k < total; k += kSinkThreads)
		a.h_ids[(base + k) % a.cap] = a.stage[(base + k) % a.cap];
}

// Several ranks: a ring slot holds one segment of spike ids per rank.  Once per window, after the
// exchange, every (step, population) list is copied into one flat list (rank order = ascending ids), so
// the delivery kernel reads spike lists the same way whatever the number of ranks.
struct flatten_args {
	std::int32_t const* const* ring_ids; // [npops]
	std::uint32_t const* const* ring_cnt;
	long long const* ring_cap;           // [npops]
	long long const* seg_lo;             // [npops][world]
	std::int32_t* const* flat_ids;       // [npops] [ring][cap]
	std::uint32_t* const* flat_cnt;      // [npops] [ring]
	int npops, ring, world;
	long long t0;
};
__global__ void __launch_bounds__(256) flatten_window(flatten_args a) {
	int const s = blockIdx.x / a.npops, p = blockIdx.x % a.npops;
	long long const slot = (a.t0 + s) % a.ring;
	std::int32_t* out    = a.flat_ids[p] + slot * a.ring_cap[p];
	unsigned done        = 0;
	for (int r = 0; r < a.world; r++) {
		unsigned const c         = a.ring_cnt[p][slot * a.world + r];
		std::int32_t const* from = a.ring_ids[p] + slot * a.ring_cap[p] + a.seg_lo[p * a.world + r];
		for (unsigned j = threadIdx.x; j < c; j += blockDim.x)
			out[done + j] = from[j];
		done += c;
	}
	if (threadIdx.x == 0)
		a.flat_cnt[p][slot] = done;
}

__global__ void fill_u32(std::uint32_t* dst, long long n, std::uint32_t value) {
	long long const i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (i < n)
		dst[i] = value;
}

// ---------------------------------------------------------------------------------------------
// host state
// ---------------------------------------------------------------------------------------------
struct population {
	spice_neuron_ops const* ops = nullptr;
	long long size = 0, lo = 0, hi = 0, stride = 0;
	std::vector<unsigned char> functor_host;
	void* functor_dev       = nullptr;
	std::uint32_t* state    = nullptr;
	std::uint64_t* history  = nullptr;
	long long rng_offset    = 0; // draws consumed per step by the populations added before this one
	std::int32_t* flat_ids  = nullptr; // world > 1: [ring][max(size, 1)] the ring's segments concatenated (flatten_window)
	std::uint32_t* flat_cnt = nullptr; // world > 1: [ring]
	// host-fed population (spice_add_host_population): spikes staged in page-locked memory, one slot per ring slot
	spice_host_update_fn host_update = nullptr;
	void* host_user                  = nullptr;
	std::int32_t* h_stage            = nullptr; // [ring][max(size, 1)]
	std::uint32_t* h_stage_cnt       = nullptr; // [ring]
	std::vector<cudaEvent_t> stage_done;        // [ring] the slot's copies have run
	u128* jump_poly         = nullptr;
	// exchange region views
	long long ring_ids_off = 0, ring_cnt_off = 0; // byte offsets in the exchange region
	std::vector<int> incoming;                     // connection indices, connect() order
	std::vector<long long> seg_lo;                 // [world]
	std::vector<long long> bounds;                 // [world + 1] rank r owns [bounds[r], bounds[r + 1]) (equal widths unless set)
};

struct connection {
	spice_synapse_ops const* ops = nullptr;
	int src = 0, dst = 0;
	long long delay = 0;
	std::vector<unsigned char> functor_host;
	void* functor_dev         = nullptr;
	long long* offsets        = nullptr;
	std::int32_t* neighbors   = nullptr;
	long long edges           = 0;
	std::uint32_t* counts     = nullptr; // [cring][cstride]
	long long cstride         = 0;       // row stride of counts: n_dst_local rounded up to 8
	apply_fn apply            = nullptr;
	// tiled delivery (deliver.cu)
	long long* tile_ptr       = nullptr; // [src][tiles + 1]
	int tile = 0, tiles = 0;
	bool duplicates           = false;   // rows may repeat a target (adj_list)
	bool arranged             = false;   // the CSR entries were replaced by the delivery stream (deliver::pack_runs)
	std::int32_t* packed      = nullptr; // arranged: 4 * groups entries
	unsigned* run_ptr         = nullptr; // arranged: [src * tiles + 1]
	// stateful / plastic synapses (window = 1 step; spice/detail/model_ops.cuh)
	bool stateful = false, plastic = false;
	bool from_to              = false;   // deliver() also reads the source neuron (concepts.h DeliverFromTo)
	std::uint32_t* src_snapshot = nullptr; // from_to: the source population's state at the end of the last step
	// from_to, world > 1: two snapshots of the WHOLE source population (by step parity) in the exchange region; every rank
	// stores its slice into every peer's copy before it publishes the step
	long long snap_off = 0, snap_half = 0, snap_stride = 0;
	std::uint32_t* syn        = nullptr; // word-SoA synapse state, parallel to neighbors
	long long syn_stride      = 0;
	std::uint64_t* ages       = nullptr; // [src] (synapse_population.h:89-94)
	std::uint32_t *evt_cnt = nullptr, *evt_off = nullptr, *evt_fill = nullptr;
	unsigned long long* evt_cursor = nullptr;
	std::int32_t* evt_list    = nullptr;
	long long evt_cap         = 0;
	apply_events_fn apply_events = nullptr;
	// fixed_probability connections are generated when the network is finalized, all of them concurrently
	bool pending_fp  = false;
	bool fp_fast     = false; // the counter-based generator (gen::generate_fixed_probability_fast): not the reference's matrix
	bool from_fp     = false; // drawn by fixed_probability (the per-synapse init hook of a sharded network regenerates the whole matrix)
	std::vector<std::int32_t> adj_src, adj_dst; // world > 1, per-synapse init hook: the adj_list's pairs until the hook has run
	double fp_p      = 0;
	UInt128 fp_seed{0, 0};   // the graph's seed (synapse_population.h:31), drawn at connect() time
	UInt128 init_seed{0, 0}; // the per-synapse init hook's (synapse_population.h:35), drawn right behind it
};

struct host_spikes {
	std::vector<std::int32_t> ids;
	long long step = -1;
};
}

struct spice_ctx {
	std::vector<long long> next_bounds; // spice_set_next_partition: the target ranges of the next population added
	int device = 0;
	cudaStream_t stream = nullptr;
	bool own_stream     = false;
	// the populations of a window update concurrently: side streams forked from / joined into `stream`
	static constexpr int kAux = 3;
	cudaStream_t aux[kAux]   = {};
	cudaEvent_t ev_fork = nullptr, ev_join[kAux] = {};
	float dt            = 0;
	long long max_delay = 1;
	util::seed_seq seed{UInt128{0, 0}};
	util::kahan_sum<float> simtime;
	long long time = 0;
	int rank = 0, world = 1, mode = 0;
	std::vector<population> pops;
	std::vector<connection> conns;
	std::string error;
	bool finalized = false;

	// geometry
	int window = 1, ring = 1, cring = 1;
	bool any_rng = false, any_stateful = false;

	// exchange region (spike rings, counters, flags)
	unsigned char* xbase = nullptr;
	size_t xbytes        = 0;
	size_t flags_off     = 0;
	unsigned char* peer_base[kMaxWorld] = {};
	bool peers_set       = false;
	unsigned long long seq = 0;

	// delivery
	bool tiled                        = true;    // SPICE_DELIVER=atomic selects the one-atomic-per-event kernel
	bool prezeroed                    = false;   // the update kernels clear the event counters they have read, so the tiled delivery may
	                                             // count a unit in several rounds that add to them (windows with few, long units)
	// Pipelined delivery: windows of at most half the shortest delay, so that the counters window w's delivery writes are
	// first read by the updates of window w + 2: delivery w runs on its own stream beside the updates of window w + 1
	// (and, with several ranks, the wait for the peers' spikes of window w is off the updates' path)
	bool pipelined                    = false;
	cudaStream_t dstream              = nullptr;
	// spike sink: sink_pack runs on its own stream beside the window's delivery and the next window's updates
	cudaStream_t sink_stream          = nullptr;
	cudaEvent_t ev_sink_ready = nullptr, ev_sink_done = nullptr;
	bool sink_pending                 = false;
	cudaEvent_t ev_upd                = nullptr;
	cudaEvent_t ev_del[2]             = {nullptr, nullptr};
	bool del_pending[2]               = {false, false};
	long long windows_run             = 0;
	bool direct_segments              = false;   // several ranks: the delivery kernel reads the ring's per-rank segments itself
	                                             // and waits for the peers' flags (no wait_window / flatten_window launches)
	deliver::conn_desc* d_conn_desc   = nullptr; // schedule order
	unsigned* d_work                  = nullptr;
	int total_tiles = 0, tile_cap = 0, n_desc = 0;

	// device tables
	std::uint32_t** d_ring_cnt        = nullptr; // [npops]
	std::uint32_t** d_peer_cnt        = nullptr; // [world][npops]
	std::int32_t** d_ring_ids         = nullptr; // [npops]
	long long* d_ring_cap             = nullptr;
	long long* d_seg_lo               = nullptr; // [npops][world]
	std::int32_t** d_flat_ids         = nullptr; // [npops] (world > 1)
	std::uint32_t** d_flat_cnt        = nullptr;
	u128* d_nib                       = nullptr;
	unsigned long long* d_stats       = nullptr; // events, spikes
	int* d_error                      = nullptr;
	long long launches                = 0;

	// spike sink (spice_raster_*)
	bool raster_on = false;
	long long sink_steps_issued = 0, sink_steps_read = 0, sink_steps_complete = 0, sink_windows = 0;
	unsigned long long sink_ids_read = 0;
	long long sink_cap = 0, sink_steps_cap = 0;
	long long* d_pop_size             = nullptr;
	unsigned long long* d_sink_cursor = nullptr; // [2]
	std::int32_t* d_sink_stage        = nullptr;
	std::int32_t* h_sink_ids          = nullptr; // page-locked, mapped
	long long* h_sink_cnt             = nullptr;
	unsigned long long* h_sink_off    = nullptr;
	unsigned long long* h_sink_consumed = nullptr; // [2] + error word behind it
	int* h_sink_error                 = nullptr;
	int sink_smem                     = 0;
	std::deque<std::pair<long long, cudaEvent_t>> sink_marks; // (steps issued up to the event, event)
	std::vector<cudaEvent_t> sink_pool;

	// readout scratch
	std::vector<std::vector<host_spikes>> spike_cache; // [pop][age]

	// phase timing (spice_profile_*): 4 events per window
	bool profile = false;
	int profile_every = 1; // phase events around every profile_every-th window (spice_profile_enable(ctx, n))
	bool profile_now  = false; // ... this one
	std::vector<cudaEvent_t> prof_events;
	size_t prof_used = 0;
	double prof_update = 0, prof_deliver = 0, prof_exchange = 0;
	long long prof_windows = 0;
};

namespace {
#define CHECK_CUDA(ctx, expr)                                                       \
	do {                                                                            \
		cudaError_t e_ = (expr);                                                    \
		if (e_ != cudaSuccess) {                                                    \
			(ctx)->error = std::string(#expr) + ": " + cudaGetErrorString(e_);     \
			return SPICE_ERR_CUDA;                                                  \
		}                                                                           \
	} while (0)

#define PRE(ctx, cond)                                                                                       \
	do {                                                                                                     \
		if (!(cond)) {                                                                                       \
			(ctx)->error = std::string("Assertion failed (") + __FILE__ + ":" + std::to_string(__LINE__) + \
			               "): " #cond;                                                                      \
			return SPICE_ERR_PRECONDITION;                                                                   \
		}                                                                                                    \
	} while (0)

int fail(spice_ctx* ctx, int code, std::string msg) {
	ctx->error = std::move(msg);
	return code;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

template <class T>
T* xptr(unsigned char* base, long long off) {
	return reinterpret_cast<T*>(base + off);
}

struct peer_blob {
	unsigned long long magic;
	int pid;
	int device;
	int rank;
	int pad;
	unsigned long long bytes;
	unsigned long long raw_ptr;
	cudaIpcMemHandle_t handle;
};

int init_synapses(spice_ctx* ctx, connection* c);

// SPICE_BUILD_TIMING=1: wall-clock of the build phases on stderr
struct build_clock {
	bool on = std::getenv("SPICE_BUILD_TIMING") != nullptr;
	std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
	void lap(char const* what) {
		if (!on)
			return;
		auto const t1 = std::chrono::steady_clock::now();
		std::fprintf(stderr, "[spice build] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
		t0 = t1;
	}
};

// Generate every pending fixed_probability connection: one host thread and one stream per connection, so the
// sequential part of each generation (one thread chasing row starts, generator.cu) overlaps the parallel parts
// of the others.
int build_pending(spice_ctx* ctx) {
	std::vector<int> todo;
	for (size_t ci = 0; ci < ctx->conns.size(); ci++)
		if (ctx->conns[ci].pending_fp)
			todo.push_back(static_cast<int>(ci));
	if (todo.empty())
		return SPICE_OK;
	CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	struct job {
		int rc = 0;
		std::string err;
		gen::result r;
	};
	std::vector<job> jobs(todo.size());
	bool const serial = std::getenv("SPICE_SERIAL_BUILD") != nullptr;
	auto run = [&](size_t j) {
		connection& c         = ctx->conns[static_cast<size_t>(todo[j])];
		population const& src = ctx->pops[c.src];
		population const& dst = ctx->pops[c.dst];
		cudaStream_t st       = nullptr;
		if (cudaSetDevice(ctx->device) != cudaSuccess || cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) {
			jobs[j].rc  = SPICE_ERR_CUDA;
			jobs[j].err = "cannot create a stream for synapse generation";
			return;
		}
		jobs[j].rc = c.fp_fast ? gen::generate_fixed_probability_fast(st, src.size, dst.size, c.fp_p, c.fp_seed.lo, c.fp_seed.hi, dst.lo, dst.hi,
		                                                              &jobs[j].r, &jobs[j].err)
		                       : gen::generate_fixed_probability(st, src.size, dst.size, c.fp_p, c.fp_seed.lo, c.fp_seed.hi, dst.lo, dst.hi, 0,
		                                                         &jobs[j].r, &jobs[j].err);
		cudaStreamDestroy(st);
	};
	if (serial || todo.size() == 1)
		for (size_t j = 0; j < todo.size(); j++)
			run(j);
	else {
		std::vector<std::thread> threads;
		for (size_t j = 0; j < todo.size(); j++)
			threads.emplace_back(run, j);
		for (auto& t : threads)
			t.join();
	}
	int rc = SPICE_OK;
	for (size_t j = 0; j < todo.size(); j++) {
		connection& c = ctx->conns[static_cast<size_t>(todo[j])];
		c.pending_fp  = false;
		if (jobs[j].rc != 0) {
			cudaFree(jobs[j].r.offsets);
			cudaFree(jobs[j].r.neighbors);
			if (rc == SPICE_OK)
				rc = fail(ctx, jobs[j].rc, jobs[j].err);
			continue;
		}
		c.offsets   = jobs[j].r.offsets;
		c.neighbors = jobs[j].r.neighbors;
		c.edges     = jobs[j].r.edges;
		ctx->launches += jobs[j].r.launches;
	}
	if (rc != SPICE_OK)
		return rc;
	for (int ci : todo) { // per-synapse state of the stateful ones (host hooks: one after the other)
		rc = init_synapses(ctx, &ctx->conns[static_cast<size_t>(ci)]);
		if (rc != SPICE_OK)
			return rc;
	}
	return SPICE_OK;
}

int finalize(spice_ctx* ctx) {
	if (ctx->finalized)
		return SPICE_OK;
	CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	build_clock clock;
	{
		int const rc = build_pending(ctx);
		if (rc != SPICE_OK)
			return rc;
	}
	clock.lap("synapse generation");
	// window = min delay over connections (any partition of the steps into windows no longer than
	// the shortest delay is valid), bounded by kMaxWindow
	long long dmin = kMaxWindow, dmax = 1;
	for (auto const& c : ctx->conns) {
		dmin = std::min(dmin, c.delay);
		dmax = std::max(dmax, c.delay);
	}
	ctx->window = static_cast<int>(std::max<long long>(1, std::min<long long>(dmin, kMaxWindow)));
	{
		// the delivery kernel stages the spike counts of (connection, step) in shared memory (deliver::kMaxCounts)
		long long stateless = 0;
		for (auto const& c : ctx->conns)
			stateless += c.stateful ? 0 : 1;
		if (stateless > 0)
			ctx->window = static_cast<int>(std::max<long long>(1, std::min<long long>(ctx->window, deliver::kMaxCounts / stateless)));
	}
	for (auto const& c : ctx->conns)
		if (c.stateful) {
			// a stateful synapse's delivery at step t depends on its target's spikes up to step t
			// (lazy plasticity) and is consumed at step t + 1: such networks advance step by step
			ctx->window       = 1;
			ctx->any_stateful = true;
		}
	{
		// SPICE_PIPELINE=1 (experiment; parity-tested, measured slower on one GPU: 2.82 against 2.70 ms per 150 steps — the
		// update kernels and the persistent delivery kernel each want whole SMs (61 K of 64 K registers, 225 of 228 KB shared
		// memory go to two delivery CTAs), so they do not run side by side, and half-length windows cost the delivery 15 %)
		char const* e  = std::getenv("SPICE_PIPELINE");
		ctx->pipelined = ctx->tiled && !ctx->any_stateful && !ctx->conns.empty() && dmin >= 2 && e && *e == '1';
		if (ctx->pipelined)
			ctx->window = static_cast<int>(std::max<long long>(1, std::min<long long>(ctx->window, dmin / 2)));
	}
	// Spike ring: a slot is read until max_delay - 1 steps after it was written (stateful delivery of a
	// connection with delay == max_delay, spikes(max_delay - 1)), and the delivery of a window reads the
	// slots that window's updates wrote.  With several ranks a peer that has passed wait_window(w) already
	// stores the spikes of window w + 1 into this rank's ring while this rank may still be reading, so the
	// ring keeps one whole window of slack behind the oldest slot anyone reads: peers are never more than
	// one window ahead (they cannot pass wait_window(w + 1) before this rank has published w + 1).
	// (pipelined: a peer may be three windows ahead of the delivery this rank is still running)
	ctx->ring   = static_cast<int>(std::max<long long>(ctx->max_delay + (ctx->world > 1 ? (ctx->pipelined ? 3 : 1) * ctx->window : 0), 2ll * ctx->window));
	ctx->cring  = static_cast<int>(std::max<long long>(dmax, 1));
	{
		size_t stateless = 0;
		for (auto const& c : ctx->conns)
			stateless += c.stateful ? 0 : 1;
		// SPICE_PREZEROED=1 (experiments): the update kernels clear the counters they read and the delivery may count a unit
		// in SPICE_DELIVER_ROUNDS rounds; measured slower than whole units at every rank shape (deliver.cu launch_tiles)
		if (char const* e = std::getenv("SPICE_PREZEROED"))
			ctx->prezeroed = *e == '1';
		ctx->direct_segments = ctx->world > 1 && !std::getenv("SPICE_FLATTEN") &&
		                       stateless * static_cast<size_t>(ctx->window) * static_cast<size_t>(ctx->world) <= static_cast<size_t>(deliver::kMaxCounts);
	}
	for (auto& p : ctx->pops)
		if (p.host_update) {
			size_t const cap = static_cast<size_t>(std::max<long long>(p.size, 1));
			CHECK_CUDA(ctx, cudaHostAlloc(&p.h_stage, sizeof(std::int32_t) * cap * ctx->ring, cudaHostAllocDefault));
			CHECK_CUDA(ctx, cudaHostAlloc(&p.h_stage_cnt, sizeof(std::uint32_t) * ctx->ring, cudaHostAllocDefault));
			p.stage_done.assign(static_cast<size_t>(ctx->ring), nullptr);
		}

	int const np = static_cast<int>(ctx->pops.size());
	if (ctx->world > 1) {
		// Load the kernels that are first launched BEHIND wait_window now: loading a kernel lazily can
		// synchronise the context, which never happens while wait_window spins for a peer whose work this
		// thread has not enqueued yet (two rank contexts in one process, tests/test_gpu_sim.py).
		cudaFuncAttributes fa{};
		CHECK_CUDA(ctx, cudaFuncGetAttributes(&fa, flatten_window));
	}
	if (ctx->world > 1 && !ctx->direct_segments) // flat copies of the spike lists for the delivery kernel (flatten_window)
		for (auto& p : ctx->pops) {
			size_t const cap = static_cast<size_t>(std::max<long long>(p.size, 1));
			CHECK_CUDA(ctx, cudaMalloc(&p.flat_ids, sizeof(std::int32_t) * cap * ctx->ring));
			CHECK_CUDA(ctx, cudaMalloc(&p.flat_cnt, sizeof(std::uint32_t) * ctx->ring));
			CHECK_CUDA(ctx, cudaMemset(p.flat_cnt, 0, sizeof(std::uint32_t) * ctx->ring));
		}
	// exchange region layout (identical on every rank)
	size_t off = 0;
	for (auto& p : ctx->pops) {
		p.ring_ids_off = static_cast<long long>(off);
		off            = align_up(off + sizeof(std::int32_t) * static_cast<size_t>(ctx->ring) * static_cast<size_t>(std::max<long long>(p.size, 1)), 256);
		p.ring_cnt_off = static_cast<long long>(off);
		off            = align_up(off + sizeof(std::uint32_t) * static_cast<size_t>(ctx->ring) * ctx->world, 256);
	}
	if (ctx->world > 1)
		for (auto& c : ctx->conns)
			if (c.from_to) {
				population const& src = ctx->pops[c.src];
				c.snap_stride         = static_cast<long long>(align_up(static_cast<size_t>(std::max<long long>(src.size, 1)), 32));
				c.snap_half           = static_cast<long long>(align_up(sizeof(std::uint32_t) * ((src.ops->neuron_bytes + 3) / 4) * static_cast<size_t>(c.snap_stride), 256));
				c.snap_off            = static_cast<long long>(off);
				off += 2 * static_cast<size_t>(c.snap_half);
			}
	ctx->flags_off = off;
	off            = align_up(off + sizeof(unsigned long long) * kMaxWorld, 256);
	ctx->xbytes    = off;
	CHECK_CUDA(ctx, cudaMalloc(&ctx->xbase, ctx->xbytes));
	CHECK_CUDA(ctx, cudaMemset(ctx->xbase, 0, ctx->xbytes));
	ctx->peer_base[ctx->rank] = ctx->xbase;

	// connections: counters + incoming lists
	{
		char const* dm = std::getenv("SPICE_DELIVER");
		ctx->tiled     = !(dm && std::string(dm) == "atomic");
	}
	long long target_len = 100; // column indices one warp-wide pass of the tiled kernel should find per row and tile
	if (char const* tl = std::getenv("SPICE_TILE_LEN"))
		target_len = std::max(8ll, std::atoll(tl));
	for (size_t ci = 0; ci < ctx->conns.size(); ci++) {
		auto& c           = ctx->conns[ci];
		auto const& src   = ctx->pops[c.src];
		auto const& dst   = ctx->pops[c.dst];
		long long const n = std::max<long long>(dst.hi - dst.lo, 1);
		c.cstride         = static_cast<long long>(align_up(static_cast<size_t>(n), 8));
		size_t const bytes = sizeof(std::uint32_t) * static_cast<size_t>(ctx->cring) * static_cast<size_t>(c.cstride);
		CHECK_CUDA(ctx, cudaMalloc(&c.counts, bytes));
		CHECK_CUDA(ctx, cudaMemset(c.counts, 0, bytes));
		ctx->pops[c.dst].incoming.push_back(static_cast<int>(ci));
		if (c.stateful) {
			c.evt_cap = std::max<long long>(1, std::min<long long>(c.edges, 1ll << 28));
			CHECK_CUDA(ctx, cudaMalloc(&c.evt_cnt, sizeof(std::uint32_t) * static_cast<size_t>(n)));
			CHECK_CUDA(ctx, cudaMalloc(&c.evt_off, sizeof(std::uint32_t) * static_cast<size_t>(n)));
			CHECK_CUDA(ctx, cudaMalloc(&c.evt_fill, sizeof(std::uint32_t) * static_cast<size_t>(n)));
			CHECK_CUDA(ctx, cudaMalloc(&c.evt_cursor, sizeof(unsigned long long)));
			CHECK_CUDA(ctx, cudaMalloc(&c.evt_list, sizeof(std::int32_t) * static_cast<size_t>(c.evt_cap)));
			CHECK_CUDA(ctx, cudaMemset(c.evt_cnt, 0, sizeof(std::uint32_t) * static_cast<size_t>(n)));
			CHECK_CUDA(ctx, cudaMemset(c.evt_cursor, 0, sizeof(unsigned long long)));
			if (c.from_to) {
				size_t const bytes = sizeof(std::uint32_t) * ((src.ops->neuron_bytes + 3) / 4) * static_cast<size_t>(src.stride);
				CHECK_CUDA(ctx, cudaMalloc(&c.src_snapshot, std::max<size_t>(bytes, 16)));
				CHECK_CUDA(ctx, cudaMemcpy(c.src_snapshot, src.state, bytes, cudaMemcpyDeviceToDevice));
			}
			if (c.plastic) {
				CHECK_CUDA(ctx, cudaMalloc(&c.ages, sizeof(std::uint64_t) * static_cast<size_t>(std::max<long long>(src.size, 1))));
				CHECK_CUDA(ctx, cudaMemset(c.ages, 0, sizeof(std::uint64_t) * static_cast<size_t>(std::max<long long>(src.size, 1))));
				// the lazy updates read the TARGET's spike history (snn.cpp:19,25); the reference
				// allocates it on the source (snn.h:46-47), which is the same population whenever
				// the reference's own behaviour is defined
				population& dp = ctx->pops[c.dst];
				if (!dp.history) {
					CHECK_CUDA(ctx, cudaMalloc(&dp.history, sizeof(std::uint64_t) * static_cast<size_t>(n)));
					CHECK_CUDA(ctx, cudaMemset(dp.history, 0, sizeof(std::uint64_t) * static_cast<size_t>(n)));
				}
			}
		}
		if (ctx->tiled && !c.stateful && c.edges > 0 && dst.hi > dst.lo) {
			// tile width: ~target_len entries of a row per tile, a multiple of 256, <= kTileMax
			double const density = static_cast<double>(c.edges) / (static_cast<double>(std::max<long long>(src.size, 1)) * static_cast<double>(n));
			long long b          = static_cast<long long>(static_cast<double>(target_len) / std::max(density, 1e-9));
			b                    = std::clamp<long long>((b + 128) / 256 * 256, 256, deliver::kTileMax);
			c.tiles              = static_cast<int>((n + b - 1) / b);
			c.tile               = static_cast<int>(std::min<long long>(b, static_cast<long long>(align_up(static_cast<size_t>((n + c.tiles - 1) / c.tiles), 256))));
			c.tiles              = static_cast<int>((n + c.tile - 1) / c.tile);
		}
	}
	if (ctx->tiled) {
		// schedule order: connections whose sources emit the most spikes per step first (largest units first)
		std::vector<int> order;
		for (size_t ci = 0; ci < ctx->conns.size(); ci++)
			if (ctx->conns[ci].tiles > 0)
				order.push_back(static_cast<int>(ci));
		std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return ctx->pops[ctx->conns[x].src].size > ctx->pops[ctx->conns[y].src].size; });
		for (int ci : order)
			ctx->tile_cap = std::max(ctx->tile_cap, static_cast<int>(align_up(static_cast<size_t>(ctx->conns[ci].tile), 128)));
		// where each tile's share of a row starts; duplicate-free connections are then rewritten into
		// the delivery kernel's own stream (bank-balanced counter addresses in whole 16-byte groups,
		// deliver.h) and give their CSR entries back
		// (one host thread and one stream per connection: the passes of different connections overlap)
		std::vector<int> pack_rc(order.size(), 0);
		std::vector<int> pack_launches(order.size(), 0);
		auto pack_one = [&](size_t j) {
			connection& c         = ctx->conns[static_cast<size_t>(order[j])];
			long long const n_src = ctx->pops[c.src].size;
			cudaStream_t st       = nullptr;
			auto check            = [&](cudaError_t e) {
                if (e != cudaSuccess && pack_rc[j] == 0)
                    pack_rc[j] = static_cast<int>(e);
                return e == cudaSuccess;
			};
			if (!check(cudaSetDevice(ctx->device)) || !check(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)))
				return;
			do {
				if (!check(cudaMalloc(&c.tile_ptr, sizeof(long long) * static_cast<size_t>(n_src) * static_cast<size_t>(c.tiles + 1))))
					break;
				if (!check(static_cast<cudaError_t>(deliver::build_tile_ptr(st, c.offsets, c.neighbors, n_src, c.tile, c.tiles, c.tile_ptr))))
					break;
				pack_launches[j]++;
				long long const n_runs = n_src * c.tiles;
				if (c.duplicates || c.edges / 4 + n_runs >= (1ll << 32)) { // multapses / more groups than 32-bit run pointers address
					check(cudaStreamSynchronize(st));
					break;
				}
				if (!check(cudaMalloc(&c.run_ptr, sizeof(unsigned) * static_cast<size_t>(n_runs + 1))))
					break;
				long long groups = 0;
				if (!check(static_cast<cudaError_t>(deliver::count_groups(st, c.tile_ptr, n_src, c.tiles, c.run_ptr, &groups))))
					break;
				if (!check(cudaMalloc(&c.packed, sizeof(std::int32_t) * 4 * static_cast<size_t>(std::max<long long>(groups, 1)))))
					break;
				if (!check(static_cast<cudaError_t>(deliver::pack_runs(st, c.neighbors, c.tile_ptr, c.run_ptr, n_src, c.tile, c.tiles, ctx->tile_cap, c.packed))))
					break;
				pack_launches[j] += 3;
				if (!check(cudaStreamSynchronize(st)))
					break;
				check(cudaFree(c.neighbors));
				check(cudaFree(c.tile_ptr));
				c.neighbors = nullptr;
				c.tile_ptr  = nullptr;
				c.arranged  = true;
			} while (false);
			cudaStreamDestroy(st);
		};
		CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		if (std::getenv("SPICE_SERIAL_BUILD") || order.size() <= 1)
			for (size_t j = 0; j < order.size(); j++)
				pack_one(j);
		else {
			std::vector<std::thread> threads;
			for (size_t j = 0; j < order.size(); j++)
				threads.emplace_back(pack_one, j);
			for (auto& t : threads)
				t.join();
		}
		for (size_t j = 0; j < order.size(); j++) {
			ctx->launches += pack_launches[j];
			CHECK_CUDA(ctx, static_cast<cudaError_t>(pack_rc[j]));
		}
		clock.lap("delivery stream (pack_runs)");
		std::vector<deliver::conn_desc> descs;
		for (int ci : order) {
			connection const& c = ctx->conns[ci];
			population const& src = ctx->pops[c.src];
			population const& dst = ctx->pops[c.dst];
			deliver::conn_desc d{};
			bool const flat = ctx->world > 1 && !ctx->direct_segments;
			d.ring_ids   = flat ? src.flat_ids : xptr<std::int32_t>(ctx->xbase, src.ring_ids_off);
			d.ring_cnt   = flat ? src.flat_cnt : xptr<std::uint32_t>(ctx->xbase, src.ring_cnt_off);
			d.ring_cap   = std::max<long long>(src.size, 1);
			d.cnt_stride = flat ? 1 : ctx->world;
			for (int r = 0; r < ctx->world; r++)
				d.seg_lo[r] = static_cast<std::int32_t>(src.bounds[static_cast<size_t>(r)]);
			d.packed      = c.packed;
			d.run_ptr     = c.run_ptr;
			d.neighbors   = c.neighbors;
			d.tile_ptr    = c.tile_ptr;
			d.counts      = c.counts;
			d.n_dst       = dst.hi - dst.lo;
			d.cstride     = c.cstride;
			d.delay       = c.delay;
			d.cring       = ctx->cring;
			d.tiles       = c.tiles;
			d.tile        = c.tile;
			d.tile_prefix = ctx->total_tiles;
			d.arranged    = c.arranged ? 1 : 0;
			ctx->total_tiles += c.tiles;
			descs.push_back(d);
		}
		if (!descs.empty()) {
			CHECK_CUDA(ctx, cudaMalloc(&ctx->d_conn_desc, sizeof(deliver::conn_desc) * descs.size()));
			CHECK_CUDA(ctx, cudaMemcpy(ctx->d_conn_desc, descs.data(), sizeof(deliver::conn_desc) * descs.size(), cudaMemcpyHostToDevice));
		}
		ctx->n_desc = static_cast<int>(descs.size());
		CHECK_CUDA(ctx, cudaMalloc(&ctx->d_work, 2 * sizeof(unsigned))); // one per window parity (pipelined delivery)
		CHECK_CUDA(ctx, cudaMemset(ctx->d_work, 0, 2 * sizeof(unsigned)));
		if (ctx->n_desc > deliver::kMaxConns)
			return fail(ctx, SPICE_ERR_UNSUPPORTED, "more than 32 stateless connections in one network");
		if (ctx->n_desc > 0)
			CHECK_CUDA(ctx, static_cast<cudaError_t>(deliver::preload()));
	}
	for (auto const& p : ctx->pops)
		if (p.incoming.size() > static_cast<size_t>(kMaxIncoming))
			return fail(ctx, SPICE_ERR_UNSUPPORTED, "more than 8 connections into one population");

	// per-population tables
	std::vector<std::uint32_t*> h_cnt(np);
	std::vector<std::int32_t*> h_ids(np);
	std::vector<long long> h_cap(np), h_seg(static_cast<size_t>(np) * ctx->world);
	long long rng_offset = 0;
	for (int pi = 0; pi < np; pi++) {
		auto& p   = ctx->pops[pi];
		h_cnt[pi] = xptr<std::uint32_t>(ctx->xbase, p.ring_cnt_off);
		h_ids[pi] = xptr<std::int32_t>(ctx->xbase, p.ring_ids_off);
		h_cap[pi] = std::max<long long>(p.size, 1);
		p.seg_lo.resize(ctx->world);
		for (int r = 0; r < ctx->world; r++) {
			p.seg_lo[r]                    = p.bounds[static_cast<size_t>(r)];
			h_seg[pi * ctx->world + r]     = p.seg_lo[r];
		}
		p.rng_offset = rng_offset;
		rng_offset += p.size * p.ops->rng_draws;
		if (p.ops->rng_draws > 0) {
			ctx->any_rng = true;
			// jump polynomial of every chunk of kRngChunk local neurons: x^(offset of its first draw)
			long long const nl     = p.hi - p.lo;
			long long const chunks = (nl + kRngChunk - 1) / kRngChunk;
			std::vector<u128> polys(static_cast<size_t>(std::max<long long>(chunks, 1)));
			util::jump::poly cur        = util::jump::xpow(static_cast<UInt>(p.rng_offset + p.lo * p.ops->rng_draws));
			util::jump::poly const step = util::jump::xpow(static_cast<UInt>(kRngChunk) * p.ops->rng_draws);
			for (long long c = 0; c < chunks; c++) {
				polys[c] = u128{cur.lo, cur.hi};
				cur      = util::jump::mulmod(cur, step);
			}
			CHECK_CUDA(ctx, cudaMalloc(&p.jump_poly, sizeof(u128) * polys.size()));
			CHECK_CUDA(ctx, cudaMemcpy(p.jump_poly, polys.data(), sizeof(u128) * polys.size(), cudaMemcpyHostToDevice));
		}
	}
	CHECK_CUDA(ctx, cudaMalloc(&ctx->d_ring_cnt, sizeof(void*) * std::max(np, 1)));
	CHECK_CUDA(ctx, cudaMalloc(&ctx->d_ring_ids, sizeof(void*) * std::max(np, 1)));
	CHECK_CUDA(ctx, cudaMalloc(&ctx->d_ring_cap, sizeof(long long) * std::max(np, 1)));
	CHECK_CUDA(ctx, cudaMalloc(&ctx->d_seg_lo, sizeof(long long) * std::max<size_t>(h_seg.size(), 1)));
	CHECK_CUDA(ctx, cudaMalloc(&ctx->d_peer_cnt, sizeof(void*) * std::max(np, 1) * ctx->world));
	if (ctx->world > 1 && !ctx->direct_segments && np) {
		std::vector<std::int32_t*> h_flat_ids;
		std::vector<std::uint32_t*> h_flat_cnt;
		for (auto const& p : ctx->pops) {
			h_flat_ids.push_back(p.flat_ids);
			h_flat_cnt.push_back(p.flat_cnt);
		}
		CHECK_CUDA(ctx, cudaMalloc(&ctx->d_flat_ids, sizeof(void*) * np));
		CHECK_CUDA(ctx, cudaMalloc(&ctx->d_flat_cnt, sizeof(void*) * np));
		CHECK_CUDA(ctx, cudaMemcpy(ctx->d_flat_ids, h_flat_ids.data(), sizeof(void*) * np, cudaMemcpyHostToDevice));
		CHECK_CUDA(ctx, cudaMemcpy(ctx->d_flat_cnt, h_flat_cnt.data(), sizeof(void*) * np, cudaMemcpyHostToDevice));
	}
	CHECK_CUDA(ctx, cudaMalloc(&ctx->d_pop_size, sizeof(long long) * std::max(np, 1)));
	if (np) {
		std::vector<long long> h_size;
		for (auto const& p : ctx->pops)
			h_size.push_back(p.size);
		CHECK_CUDA(ctx, cudaMemcpy(ctx->d_pop_size, h_size.data(), sizeof(long long) * np, cudaMemcpyHostToDevice));
		CHECK_CUDA(ctx, cudaMemcpy(ctx->d_ring_cnt, h_cnt.data(), sizeof(void*) * np, cudaMemcpyHostToDevice));
		CHECK_CUDA(ctx, cudaMemcpy(ctx->d_ring_ids, h_ids.data(), sizeof(void*) * np, cudaMemcpyHostToDevice));
		CHECK_CUDA(ctx, cudaMemcpy(ctx->d_ring_cap, h_cap.data(), sizeof(long long) * np, cudaMemcpyHostToDevice));
		CHECK_CUDA(ctx, cudaMemcpy(ctx->d_seg_lo, h_seg.data(), sizeof(long long) * h_seg.size(), cudaMemcpyHostToDevice));
	}
	if (ctx->any_rng)
		CHECK_CUDA(ctx, cudaMalloc(&ctx->d_nib, sizeof(u128) * kMaxWindow * 32 * 16));
	CHECK_CUDA(ctx, cudaMalloc(&ctx->d_stats, sizeof(unsigned long long) * 2));
	CHECK_CUDA(ctx, cudaMemset(ctx->d_stats, 0, sizeof(unsigned long long) * 2));
	CHECK_CUDA(ctx, cudaMalloc(&ctx->d_error, sizeof(int)));
	CHECK_CUDA(ctx, cudaMemset(ctx->d_error, 0, sizeof(int)));
	ctx->spike_cache.assign(np, std::vector<host_spikes>(static_cast<size_t>(ctx->ring)));
	clock.lap("tables");
	ctx->finalized = true;
	return SPICE_OK;
}

int upload_peer_tables(spice_ctx* ctx) {
	int const np = static_cast<int>(ctx->pops.size());
	std::vector<std::uint32_t*> h(static_cast<size_t>(std::max(np, 1)) * ctx->world);
	for (int r = 0; r < ctx->world; r++)
		for (int pi = 0; pi < np; pi++)
			h[r * np + pi] = xptr<std::uint32_t>(ctx->peer_base[r], ctx->pops[pi].ring_cnt_off);
	if (np)
		CHECK_CUDA(ctx, cudaMemcpy(ctx->d_peer_cnt, h.data(), sizeof(void*) * np * ctx->world, cudaMemcpyHostToDevice));
	return SPICE_OK;
}

void fill_incoming(spice_ctx* ctx, population const& p, incoming* in, int* n_in) {
	*n_in = static_cast<int>(p.incoming.size());
	for (int k = 0; k < *n_in; k++) {
		connection const& c = ctx->conns[p.incoming[k]];
		in[k]               = incoming{c.counts, c.cstride, c.apply, c.functor_dev, ctx->cring, (ctx->tiled && !ctx->prezeroed) ? 0 : 1,
		                               c.stateful ? c.evt_cnt : nullptr, c.evt_off, c.evt_list, c.syn, c.syn_stride, c.apply_events,
		                               from_ctx{c.src_snapshot, ctx->pops[c.src].stride, reinterpret_cast<std::int64_t const*>(c.offsets),
		                                        ctx->pops[c.src].size, ctx->mode == SPICE_MODE_FAST ? 1 : 0}};
		if (c.from_to && ctx->world > 1) { // the whole source population as every rank left it at the end of the step before
			in[k].from.state  = xptr<std::uint32_t>(ctx->xbase, c.snap_off + ((ctx->time + 1) & 1) * c.snap_half);
			in[k].from.stride = c.snap_stride;
		}
	}
}

cudaEvent_t prof_mark(spice_ctx* ctx, cudaStream_t on = nullptr) {
	if (ctx->prof_used == ctx->prof_events.size()) {
		cudaEvent_t e = nullptr;
		cudaEventCreate(&e);
		ctx->prof_events.push_back(e);
	}
	cudaEvent_t e = ctx->prof_events[ctx->prof_used++];
	cudaEventRecord(e, on ? on : ctx->stream);
	return e;
}

int prof_collect(spice_ctx* ctx) {
	if (ctx->prof_used == 0)
		return SPICE_OK;
	CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	if (ctx->dstream)
		CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->dstream));
	for (size_t i = 0; i + 3 < ctx->prof_used; i += 4) {
		float a = 0, b = 0, c = 0;
		cudaEventElapsedTime(&a, ctx->prof_events[i], ctx->prof_events[i + 1]);
		cudaEventElapsedTime(&b, ctx->prof_events[i + 1], ctx->prof_events[i + 2]);
		cudaEventElapsedTime(&c, ctx->prof_events[i + 2], ctx->prof_events[i + 3]);
		ctx->prof_update += a;
		ctx->prof_exchange += b;
		ctx->prof_deliver += c;
		ctx->prof_windows++;
	}
	ctx->prof_used = 0;
	return SPICE_OK;
}

int run_window(spice_ctx* ctx, int nsteps) {
	int const np = static_cast<int>(ctx->pops.size());
	ctx->profile_now = ctx->profile && ctx->windows_run % ctx->profile_every == 0;
	if (ctx->profile_now) {
		if (ctx->prof_used >= 4096) {
			int const rc = prof_collect(ctx);
			if (rc != SPICE_OK)
				return rc;
		}
		prof_mark(ctx);
	}
	// host-side per-step scalars: compensated dt (snn.cpp:8-10) and the step's stream seed (snn.cpp:12)
	prologue_args pa{};
	float dts[kMaxWindow];
	for (int s = 0; s < nsteps; s++) {
		dts[s] = ctx->simtime += ctx->dt;
		if (ctx->simtime >= 1)
			ctx->simtime.reset();
		UInt128 const sd = (ctx->seed++).seed();
		pa.seed[s]       = u128{sd.lo, sd.hi};
	}
	pa.nib      = ctx->any_rng ? ctx->d_nib : nullptr;
	pa.ring_cnt = ctx->d_ring_cnt;
	pa.npops    = np;
	pa.ring     = ctx->ring;
	pa.world    = ctx->world;
	pa.rank     = ctx->rank;
	pa.t0       = ctx->time;
	pa.nsteps   = nsteps;
	int const parity = static_cast<int>(ctx->windows_run & 1);
	if (ctx->pipelined) {
		if (!ctx->dstream) {
			CHECK_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->dstream, cudaStreamNonBlocking));
			CHECK_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_upd, cudaEventDisableTiming));
			for (auto& e : ctx->ev_del)
				CHECK_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
		}
		// this window's updates read the counters delivery w - 2 wrote (and reuse its work counter); delivery w - 1 may
		// still be running: its counters are first read by the next window
		if (ctx->del_pending[parity])
			CHECK_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_del[parity], 0));
	}
	pa.work     = ctx->d_work + (ctx->pipelined ? parity : 0);
	window_prologue<<<nsteps, 512, 0, ctx->stream>>>(pa);
	ctx->launches++;

	// host-fed populations: ask the host for every step's spikes, in step order, and copy them into the ring
	for (int s = 0; s < nsteps; s++) {
		long long host_draws = 0; // drawn by this step's host functors so far: the next one's stream position moves by that much
		for (auto& p : ctx->pops) {
			if (!p.host_update) {
				// a device population jumps into the step's stream at a position fixed when the network was built
				if (host_draws != 0 && p.ops->rng_draws != 0)
					return fail(ctx, SPICE_ERR_UNSUPPORTED,
					            "a per-population update() that draws from the step's random stream must come after every device population that draws");
				continue;
			}
			long long const slot = (ctx->time + s) % ctx->ring;
			long long const cap  = std::max<long long>(p.size, 1);
			cudaEvent_t& done    = p.stage_done[static_cast<size_t>(slot)];
			if (done)
				CHECK_CUDA(ctx, cudaEventSynchronize(done)); // the slot's previous copies have run
			else
				CHECK_CUDA(ctx, cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
			int64_t draws   = 0;
			int64_t const n = p.host_update(p.host_user, dts[s], pa.seed[s].lo, pa.seed[s].hi, static_cast<uint64_t>(p.rng_offset + host_draws),
			                                p.h_stage + slot * cap, p.size, &draws);
			if (n < 0 || n > p.size || draws < 0)
				return fail(ctx, SPICE_ERR_PRECONDITION, "per-population update(): spike count out of range");
			host_draws += draws;
			for (int64_t i = 0; i < n; i++)
				if (p.h_stage[slot * cap + i] < 0 || p.h_stage[slot * cap + i] >= p.size)
					return fail(ctx, SPICE_ERR_PRECONDITION, "per-population update(): spike id out of range");
			p.h_stage_cnt[slot] = static_cast<std::uint32_t>(n);
			if (n)
				CHECK_CUDA(ctx, cudaMemcpyAsync(xptr<std::int32_t>(ctx->xbase, p.ring_ids_off) + slot * cap, p.h_stage + slot * cap,
				                                sizeof(std::int32_t) * static_cast<size_t>(n), cudaMemcpyHostToDevice, ctx->stream));
			CHECK_CUDA(ctx, cudaMemcpyAsync(xptr<std::uint32_t>(ctx->xbase, p.ring_cnt_off) + slot * ctx->world, p.h_stage_cnt + slot,
			                                sizeof(std::uint32_t), cudaMemcpyHostToDevice, ctx->stream));
			CHECK_CUDA(ctx, cudaEventRecord(done, ctx->stream));
		}
	}

	// The populations' update kernels are independent of each other (each reads counters written in
	// earlier windows and writes its own state and spike lists), and the smaller ones do not fill the
	// device: all but the first run on side streams, forked here and joined before the exchange.
	int device_pops = 0;
	for (auto const& p : ctx->pops)
		device_pops += (!p.host_update && p.hi > p.lo) ? 1 : 0;
	bool const fan = device_pops > 1 && !std::getenv("SPICE_SERIAL_UPDATES");
	if (fan) {
		if (!ctx->ev_fork) {
			CHECK_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
			for (int k = 0; k < spice_ctx::kAux; k++) {
				CHECK_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->aux[k], cudaStreamNonBlocking));
				CHECK_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_join[k], cudaEventDisableTiming));
			}
		}
		CHECK_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
	}
	int launched = 0;
	unsigned aux_used = 0;
	for (int pi = 0; pi < np; pi++) {
		population& p = ctx->pops[pi];
		if (p.host_update)
			continue;
		update_args ua{};
		ua.stream  = ctx->stream;
		if (fan && p.hi > p.lo && launched++ > 0) {
			int const k = (launched - 2) % spice_ctx::kAux;
			if (!((aux_used >> k) & 1u)) {
				CHECK_CUDA(ctx, cudaStreamWaitEvent(ctx->aux[k], ctx->ev_fork, 0));
				aux_used |= 1u << k;
			}
			ua.stream = ctx->aux[k];
		}
		ua.functor = p.functor_dev;
		ua.state   = p.state;
		ua.n_local = p.hi - p.lo;
		ua.lo      = p.lo;
		ua.stride  = p.stride;
		ua.t0      = ctx->time;
		ua.nsteps  = nsteps;
		ua.cring   = ctx->cring;
		ua.cslot0  = static_cast<int>(ctx->time % ctx->cring);
		ua.rslot0  = static_cast<int>(ctx->time % ctx->ring);
		std::memcpy(ua.dt, dts, sizeof(float) * nsteps);
		for (int r = 0; r < ctx->world; r++)
			ua.ring_ids[r] = xptr<std::int32_t>(ctx->peer_base[r], p.ring_ids_off);
		ua.ring_cnt = xptr<std::uint32_t>(ctx->xbase, p.ring_cnt_off);
		ua.ring_cap = std::max<long long>(p.size, 1);
		ua.ring     = ctx->ring;
		ua.rank     = ctx->rank;
		ua.world    = ctx->world;
		ua.history  = p.history;
		fill_incoming(ctx, p, ua.in, &ua.n_in);
		ua.error     = ctx->d_error;
		ua.rng.nib   = ctx->d_nib;
		ua.jump_poly = p.jump_poly;
		if (ua.n_local > 0) {
			// all incoming connections stateless, fed by the tiled delivery and of one synapse type: the
			// update kernel with that type's deliver() inlined
			int (*fused)(update_args const*) = nullptr;
			if (ua.n_in >= 1 && ua.n_in <= 4 && ctx->tiled && !std::getenv("SPICE_UPDATE_GENERIC")) {
				fused = ctx->conns[p.incoming[0]].ops->launch_update_fused;
				for (int ci : p.incoming) {
					connection const& c = ctx->conns[ci];
					if (c.stateful || c.ops->launch_update_fused != fused)
						fused = nullptr;
				}
				for (int c = 0; c < ua.n_in; c++)
					if (ua.in[c].evt_cnt)
						fused = nullptr;
			}
			int const e = fused ? fused(&ua) : p.ops->launch_update(&ua);
			if (e != 0)
				return fail(ctx, SPICE_ERR_CUDA, std::string("update launch: ") + cudaGetErrorString(static_cast<cudaError_t>(e)));
			ctx->launches++;
		}
	}
	for (int k = 0; k < spice_ctx::kAux; k++)
		if ((aux_used >> k) & 1u) {
			CHECK_CUDA(ctx, cudaEventRecord(ctx->ev_join[k], ctx->aux[k]));
			CHECK_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join[k], 0));
		}

	// the sink of the window before reads ring slots that the NEXT window (this rank's prologue, the peers' updates once
	// they have this window's flag) writes again: it has had that window's delivery and this window's updates to finish
	if (ctx->sink_pending) {
		CHECK_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_sink_done, 0));
		ctx->sink_pending = false;
	}
	if (ctx->profile_now)
		prof_mark(ctx);
	if (ctx->world > 1) {
		// DeliverFromTo synapses read their source neuron as it is at the end of this step, wherever it lives: this rank's
		// slice of the source population goes into every rank's snapshot of parity (step & 1) ahead of the window flag.  A peer
		// is at most one step ahead (it needs this rank's flag of step t to start step t + 2), so two snapshots suffice.
		for (auto& c : ctx->conns)
			if (c.from_to) {
				population const& src = ctx->pops[c.src];
				if (src.hi <= src.lo)
					continue;
				size_t const words = (src.ops->neuron_bytes + 3) / 4;
				for (int r = 0; r < ctx->world; r++)
					CHECK_CUDA(ctx, cudaMemcpy2DAsync(xptr<std::uint32_t>(ctx->peer_base[r], c.snap_off + (ctx->time & 1) * c.snap_half) + src.lo,
					                                  sizeof(std::uint32_t) * static_cast<size_t>(c.snap_stride), src.state,
					                                  sizeof(std::uint32_t) * static_cast<size_t>(src.stride),
					                                  sizeof(std::uint32_t) * static_cast<size_t>(src.hi - src.lo), words, cudaMemcpyDeviceToDevice,
					                                  ctx->stream));
			}
		ctx->seq++;
		publish_args pb{};
		pb.local_cnt = ctx->d_ring_cnt;
		pb.peer_cnt  = ctx->d_peer_cnt;
		for (int r = 0; r < ctx->world; r++)
			pb.peer_flags[r] = xptr<unsigned long long>(ctx->peer_base[r], static_cast<long long>(ctx->flags_off));
		pb.npops  = np;
		pb.ring   = ctx->ring;
		pb.world  = ctx->world;
		pb.rank   = ctx->rank;
		pb.t0     = ctx->time;
		pb.nsteps = nsteps;
		pb.seq    = ctx->seq;
		publish_window<<<1, 256, 0, ctx->stream>>>(pb);
		ctx->launches++;
		// who waits for the peers: the tiled delivery launch itself when it reads the ring's segments directly and the
		// window has nothing else to run behind the exchange (stateful deliveries, raster packing); else a wait kernel
		bool const deliver_waits = ctx->direct_segments && ctx->tiled && ctx->n_desc > 0 && !ctx->any_stateful && !ctx->raster_on;
		if (!deliver_waits) {
			wait_args wa{xptr<unsigned long long>(ctx->xbase, static_cast<long long>(ctx->flags_off)), ctx->world, ctx->seq, ctx->d_error};
			wait_window<<<1, 32, 0, ctx->stream>>>(wa);
			ctx->launches++;
		}
		if (ctx->tiled && np > 0 && !ctx->direct_segments) { // one flat spike list per (step, population) for the delivery kernel
			flatten_args fa{ctx->d_ring_ids, ctx->d_ring_cnt, ctx->d_ring_cap, ctx->d_seg_lo, ctx->d_flat_ids, ctx->d_flat_cnt,
			                np,              ctx->ring,       ctx->world,      ctx->time};
			flatten_window<<<nsteps * np, 256, 0, ctx->stream>>>(fa);
			ctx->launches++;
		}
	}
	// every spike of the window is in this rank's ring now: the sink may read it from here on (on its own stream)
	bool const sink_aside = ctx->raster_on && np > 0 && !ctx->pipelined && !std::getenv("SPICE_SINK_INLINE");
	if (sink_aside) {
		if (!ctx->sink_stream) {
			CHECK_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->sink_stream, cudaStreamNonBlocking));
			CHECK_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_sink_ready, cudaEventDisableTiming));
			CHECK_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_sink_done, cudaEventDisableTiming));
		}
		CHECK_CUDA(ctx, cudaEventRecord(ctx->ev_sink_ready, ctx->stream));
	}

	if (ctx->profile_now)
		prof_mark(ctx);
	if (ctx->any_stateful) {
		// nsteps == 1 here.  snn.cpp:17-25: every 64 steps catch every plastic synapse up, then deliver
		// the spikes emitted delay - 1 steps ago, connection by connection
		for (int pass = 0; pass < 2; pass++)
			for (auto& c : ctx->conns) {
				if (!c.stateful)
					continue;
				population const& src = ctx->pops[c.src];
				population const& dst = ctx->pops[c.dst];
				if (dst.hi - dst.lo <= 0)
					continue;
				if (pass == 0 && !(c.plastic && ctx->time % 64 == 0))
					continue;
				if (pass == 1 && ctx->time < c.delay - 1)
					continue;
				long long const slot = (ctx->time - (c.delay - 1) + ctx->ring) % ctx->ring;
				stateful_args sa{};
				sa.stream   = ctx->stream;
				sa.functor  = c.functor_dev;
				sa.ring_ids = xptr<std::int32_t>(ctx->xbase, src.ring_ids_off) + slot * std::max<long long>(src.size, 1);
				sa.ring_cnt = xptr<std::uint32_t>(ctx->xbase, src.ring_cnt_off) + slot * ctx->world;
				for (int r = 0; r < ctx->world; r++)
					sa.seg_lo[r] = src.bounds[static_cast<size_t>(r)];
				sa.world       = ctx->world;
				sa.n_src       = src.size;
				sa.n_dst       = dst.hi - dst.lo;
				sa.offsets     = reinterpret_cast<std::int64_t const*>(c.offsets);
				sa.neighbors   = c.neighbors;
				sa.syn         = c.syn;
				sa.syn_stride  = c.syn_stride;
				sa.dst_history = dst.history;
				sa.ages        = c.ages;
				sa.time        = ctx->time;
				sa.dt          = ctx->dt;
				sa.evt_cnt     = c.evt_cnt;
				sa.evt_off     = c.evt_off;
				sa.evt_fill    = c.evt_fill;
				sa.evt_cursor  = c.evt_cursor;
				sa.evt_list    = c.evt_list;
				sa.evt_cap     = c.evt_cap;
				sa.stats       = ctx->d_stats;
				sa.error       = ctx->d_error;
				for (int phase : {pass == 0 ? 3 : 0, 1, 2}) {
					sa.phase    = phase;
					int const e = c.ops->launch_stateful(&sa);
					if (e != 0)
						return fail(ctx, SPICE_ERR_CUDA, std::string("stateful delivery launch: ") + cudaGetErrorString(static_cast<cudaError_t>(e)));
					ctx->launches++;
					if (pass == 0)
						break;
				}
			}
		// deliver() of a DeliverFromTo synapse sees its source as it is now, at the end of the step
		// (synapse_population.h:125-131 runs after every population's update, snn.cpp:21-25); the events
		// counted above are applied at the start of the next step, so keep that state
		for (auto& c : ctx->conns)
			if (c.from_to && ctx->world == 1) { // (more than one rank: stored into every rank's snapshot before the exchange)
				population const& src = ctx->pops[c.src];
				size_t const bytes    = sizeof(std::uint32_t) * ((src.ops->neuron_bytes + 3) / 4) * static_cast<size_t>(src.stride);
				CHECK_CUDA(ctx, cudaMemcpyAsync(c.src_snapshot, src.state, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
			}
	}
	if (ctx->tiled) {
		if (ctx->n_desc > 0) {
			deliver::tiles_args ta{};
			ta.conns       = ctx->d_conn_desc;
			ta.nconns      = ctx->n_desc;
			ta.total_tiles = ctx->total_tiles;
			ta.ring        = ctx->ring;
			ta.world       = ctx->world;
			ta.t0          = ctx->time;
			ta.nsteps      = nsteps;
			ta.work        = ctx->d_work;
			ta.stats       = ctx->d_stats;
			ta.error       = ctx->d_error;
			ta.tile_cap    = ctx->tile_cap;
			ta.prezeroed   = ctx->prezeroed ? 1 : 0;
			ta.world       = ctx->world > 1 && !ctx->direct_segments ? 1 : ctx->world; // flat lists look like one rank's
			if (ctx->direct_segments && !ctx->any_stateful && !ctx->raster_on) {
				ta.flags = xptr<unsigned long long>(ctx->xbase, static_cast<long long>(ctx->flags_off));
				ta.seq   = ctx->seq;
			}
			int launched   = 0;
			cudaStream_t on = ctx->stream;
			if (ctx->pipelined) { // beside the next window's updates: after this window's updates (and publish), before the updates of w + 2
				on      = ctx->dstream;
				ta.work = ctx->d_work + parity;
				CHECK_CUDA(ctx, cudaEventRecord(ctx->ev_upd, ctx->stream));
				CHECK_CUDA(ctx, cudaStreamWaitEvent(on, ctx->ev_upd, 0));
				if (ctx->profile_now) {
					ctx->prof_used--; // the mark behind the exchange goes where the delivery starts
					prof_mark(ctx, on);
				}
			}
			int const e    = deliver::launch_tiles(on, ta, ctx->device, &launched);
			if (e != 0)
				return fail(ctx, SPICE_ERR_CUDA, std::string("delivery launch: ") + cudaGetErrorString(static_cast<cudaError_t>(e)));
			ctx->launches += launched;
			if (ctx->pipelined) {
				CHECK_CUDA(ctx, cudaEventRecord(ctx->ev_del[parity], on));
				ctx->del_pending[parity] = true;
			}
		}
	} else
		for (auto& c : ctx->conns) {
			population const& src = ctx->pops[c.src];
			population const& dst = ctx->pops[c.dst];
			if (dst.hi - dst.lo <= 0 || c.edges == 0 || c.stateful)
				continue;
			deliver_args da{};
			da.ring_ids = xptr<std::int32_t>(ctx->xbase, src.ring_ids_off);
			da.ring_cnt = xptr<std::uint32_t>(ctx->xbase, src.ring_cnt_off);
			da.ring_cap = std::max<long long>(src.size, 1);
			da.ring     = ctx->ring;
			da.world    = ctx->world;
			for (int r = 0; r < ctx->world; r++)
				da.seg_lo[r] = src.seg_lo[r];
			da.offsets     = c.offsets;
			da.neighbors   = c.neighbors;
			da.counts      = c.counts;
			da.cstride     = c.cstride;
			da.cring       = ctx->cring;
			da.delay       = c.delay;
			da.t0          = ctx->time;
			da.nsteps      = nsteps;
			da.stats       = ctx->d_stats;
			// enough warps to cover the window's spikes at typical rates; the kernel strides
			long long const est = std::max<long long>(1, src.size / 64) * nsteps;
			int const blocks    = static_cast<int>(std::min<long long>(148 * 8, (est + 7) / 8));
			deliver_counts<<<std::max(blocks, 1), 256, 0, ctx->stream>>>(da);
			ctx->launches++;
		}

	if (ctx->profile_now)
		prof_mark(ctx, ctx->pipelined && ctx->tiled && ctx->n_desc > 0 ? ctx->dstream : nullptr);
	if (ctx->raster_on && np > 0) {
		sink_args sa{};
		sa.ring_ids    = ctx->d_ring_ids;
		sa.ring_cnt    = ctx->d_ring_cnt;
		sa.ring_cap    = ctx->d_ring_cap;
		sa.seg_lo      = ctx->d_seg_lo;
		sa.pop_size    = ctx->d_pop_size;
		sa.npops       = np;
		sa.ring        = ctx->ring;
		sa.world       = ctx->world;
		sa.t0          = ctx->time;
		sa.nsteps      = nsteps;
		sa.parity      = static_cast<int>(ctx->sink_windows & 1);
		sa.step_index0 = ctx->sink_steps_issued;
		sa.cursor      = ctx->d_sink_cursor;
		sa.stage       = ctx->d_sink_stage;
		sa.h_ids       = ctx->h_sink_ids;
		sa.cap         = ctx->sink_cap;
		sa.h_cnt       = ctx->h_sink_cnt;
		sa.h_off       = ctx->h_sink_off;
		sa.steps_cap   = ctx->sink_steps_cap;
		sa.consumed_steps = static_cast<unsigned long long>(ctx->sink_steps_read);
		sa.consumed_ids   = ctx->sink_ids_read;
		sa.h_error     = ctx->h_sink_error;
		cudaStream_t const sink_on = sink_aside ? ctx->sink_stream : ctx->stream;
		if (sink_aside)
			CHECK_CUDA(ctx, cudaStreamWaitEvent(sink_on, ctx->ev_sink_ready, 0));
		sink_pack<<<nsteps * np, kSinkThreads, ctx->sink_smem, sink_on>>>(sa);
		ctx->launches++;
		if (sink_aside) {
			CHECK_CUDA(ctx, cudaEventRecord(ctx->ev_sink_done, sink_on));
			ctx->sink_pending = true;
		}
		ctx->sink_windows++;
		ctx->sink_steps_issued += nsteps;
		// completion mark of this window: a readout waits for the event of the steps it takes, not for the stream
		cudaEvent_t ev = nullptr;
		if (!ctx->sink_pool.empty()) {
			ev = ctx->sink_pool.back();
			ctx->sink_pool.pop_back();
		} else if (ctx->sink_marks.size() >= 4096) { // nobody is reading: keep the newest marks (a later event covers earlier steps)
			ev = ctx->sink_marks.front().second;
			ctx->sink_marks.pop_front();
		} else if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess)
			return fail(ctx, SPICE_ERR_CUDA, "cudaEventCreate (spike sink)");
		if (cudaEventRecord(ev, sink_on) != cudaSuccess)
			return fail(ctx, SPICE_ERR_CUDA, "cudaEventRecord (spike sink)");
		ctx->sink_marks.emplace_back(ctx->sink_steps_issued, ev);
	}
	cudaError_t const e = cudaGetLastError();
	if (e != cudaSuccess)
		return fail(ctx, SPICE_ERR_CUDA, std::string("window launch: ") + cudaGetErrorString(e));
	ctx->time += nsteps;
	ctx->windows_run++;
	return SPICE_OK;
}

int check_device_error(spice_ctx* ctx) {
	int h = 0;
	CHECK_CUDA(ctx, cudaMemcpyAsync(&h, ctx->d_error, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
	CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	if (h & 1)
		return fail(ctx, SPICE_ERR_INTERNAL, "spike exchange timed out waiting for a peer rank");
	if (h & 16)
		return fail(ctx, SPICE_ERR_INTERNAL, "spike delivery: internal error (pipeline made no progress)");
	if (h & 32)
		return fail(ctx, SPICE_ERR_INTERNAL, "stateful delivery: event list capacity exceeded");
	if (h & 64)
		return fail(ctx, SPICE_ERR_PRECONDITION, "a neuron's update() drew more random numbers than its rng_draws declares");
	if (h & 4)
		return fail(ctx, SPICE_ERR_INTERNAL, "raster log: step capacity exceeded (read the raster more often)");
	if (h & 8)
		return fail(ctx, SPICE_ERR_INTERNAL, "raster log: id capacity exceeded (read the raster more often)");
	return SPICE_OK;
}

int add_connection_common(spice_ctx* ctx, spice_synapse_ops const* ops, int src_pop, int dst_pop, float delay,
                          void const* functor, connection* c) {
	PRE(ctx, ops != nullptr && ops->abi_version == 1);
	PRE(ctx, !ctx->finalized && "connect() after the first step is not supported");
	PRE(ctx, src_pop >= 0 && src_pop < static_cast<int>(ctx->pops.size()));
	PRE(ctx, dst_pop >= 0 && dst_pop < static_cast<int>(ctx->pops.size()));
	PRE(ctx, ctx->pops[dst_pop].ops->neuron_bytes == ops->dst_neuron_bytes);
	long long const d = static_cast<long long>(std::round(delay / ctx->dt)); // snn.h:33
	PRE(ctx, d >= 1 && "The delay must be at least 1dt.");                 // snn.h:35
	PRE(ctx, d <= ctx->max_delay && "The delay of a synapse population may not exceed the maximum delay of the network."); // snn.h:36-38
	if (ops->deliver_from_to && ops->synapse_bytes == 0)
		return fail(ctx, SPICE_ERR_UNSUPPORTED, // (the facade hands stateless ones over with one carried word of state: model_ops.cuh carried_from_to)
		            "a synapse whose deliver() reads the source neuron needs per-synapse state in this ABI");
	PRE(ctx, ops->synapse_bytes % 4 == 0);
	c->stateful = ops->synapse_bytes != 0;
	c->from_to  = ops->deliver_from_to != 0;
	c->plastic  = ops->plastic != 0;
	c->ops   = ops;
	c->src   = src_pop;
	c->dst   = dst_pop;
	c->delay = d;
	c->functor_host.assign(static_cast<unsigned char const*>(functor), static_cast<unsigned char const*>(functor) + ops->functor_bytes);
	CHECK_CUDA(ctx, cudaMalloc(&c->functor_dev, std::max<size_t>(ops->functor_bytes, 16)));
	CHECK_CUDA(ctx, cudaMemcpy(c->functor_dev, functor, ops->functor_bytes, cudaMemcpyHostToDevice));
	int e = ops->get_apply(&c->apply);
	if (e == 0 && c->stateful)
		e = ops->get_apply_events(&c->apply_events);
	if (e != 0) {
		cudaFree(c->functor_dev);
		c->functor_dev = nullptr;
		return fail(ctx, SPICE_ERR_CUDA, std::string("get_apply: ") + cudaGetErrorString(static_cast<cudaError_t>(e)));
	}
	return SPICE_OK;
}

// a connection that failed before it reached ctx->conns (whose entries spice_ctx_destroy frees)
void drop_connection(connection& c) {
	cudaFree(c.functor_dev);
	cudaFree(c.offsets);
	cudaFree(c.neighbors);
	cudaFree(c.syn);
	c.functor_dev = nullptr;
	c.offsets     = nullptr;
	c.neighbors   = nullptr;
	c.syn         = nullptr;
}

// per-synapse state of a stateful connection: default-constructed synapses, then the model's init
// hook (synapse_population.h:34-41), stored word-SoA parallel to the CSR
int init_synapses(spice_ctx* ctx, connection* c) {
	if (!c->stateful)
		return SPICE_OK;
	PRE(ctx, c->edges < 2147483647ll && "a stateful connection holds at most 2^31 - 1 synapses per rank");
	int const words       = static_cast<int>(c->ops->synapse_bytes / 4);
	population const& src = ctx->pops[c->src];
	population const& dst = ctx->pops[c->dst];
	c->syn_stride         = static_cast<long long>(align_up(static_cast<size_t>(std::max<long long>(c->edges, 1)), 32));
	CHECK_CUDA(ctx, cudaMalloc(&c->syn, sizeof(std::uint32_t) * static_cast<size_t>(words) * static_cast<size_t>(c->syn_stride)));
	if (c->ops->per_synapse_init && ctx->world > 1) {
		// The hook walks the WHOLE connection with one engine (synapse_population.h:34-41), so what a synapse receives
		// depends on every synapse before it: each rank builds the whole adjacency once more (all columns, a temporary),
		// runs the hook over it on the host and keeps the states of its own columns.  Costs the full matrix in host memory
		// per rank; hooks are a feature of small networks (samples/ping_pong, sssp).
		UInt128 const sd = c->init_seed;
		gen::result full;
		std::string err;
		int grc = 0;
		CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		if (c->from_fp)
			grc = c->fp_fast ? gen::generate_fixed_probability_fast(ctx->stream, src.size, dst.size, c->fp_p, c->fp_seed.lo, c->fp_seed.hi, 0, dst.size, &full, &err)
			                 : gen::generate_fixed_probability(ctx->stream, src.size, dst.size, c->fp_p, c->fp_seed.lo, c->fp_seed.hi, 0, dst.size, 0, &full, &err);
		else {
			bool dup = false;
			grc      = gen::generate_adj_list(ctx->stream, c->adj_src.data(), c->adj_dst.data(), static_cast<long long>(c->adj_src.size()), src.size,
			                                  dst.size, 0, dst.size, &full, &dup, &err);
		}
		std::vector<long long> off(static_cast<size_t>(src.size) + 1);
		std::vector<std::int32_t> nb(static_cast<size_t>(std::max<long long>(full.edges, 1)));
		cudaError_t e1 = cudaSuccess, e2 = cudaSuccess;
		if (grc == 0) {
			e1 = cudaMemcpy(off.data(), full.offsets, sizeof(long long) * off.size(), cudaMemcpyDeviceToHost);
			if (full.edges)
				e2 = cudaMemcpy(nb.data(), full.neighbors, sizeof(std::int32_t) * static_cast<size_t>(full.edges), cudaMemcpyDeviceToHost);
		}
		cudaFree(full.offsets);
		cudaFree(full.neighbors);
		if (grc != 0)
			return fail(ctx, grc == 1 ? SPICE_ERR_PRECONDITION : grc, err);
		CHECK_CUDA(ctx, e1);
		CHECK_CUDA(ctx, e2);
		std::vector<unsigned char> aos(static_cast<size_t>(std::max<long long>(full.edges, 1)) * c->ops->synapse_bytes);
		c->ops->init_host(c->functor_host.data(), aos.data(), reinterpret_cast<std::int64_t const*>(off.data()), nb.data(), src.size, sd.lo, sd.hi);
		std::vector<std::uint32_t> soa(static_cast<size_t>(words) * static_cast<size_t>(c->syn_stride), 0);
		long long mine = 0; // this rank's synapses are the whole matrix's entries with a local target, in the same order
		for (long long e = 0; e < full.edges; e++) {
			if (nb[static_cast<size_t>(e)] < dst.lo || nb[static_cast<size_t>(e)] >= dst.hi)
				continue;
			if (mine < c->edges)
				for (int w = 0; w < words; w++)
					std::memcpy(&soa[static_cast<size_t>(w) * c->syn_stride + mine], aos.data() + e * c->ops->synapse_bytes + 4 * w, 4);
			mine++;
		}
		if (mine != c->edges)
			return fail(ctx, SPICE_ERR_INTERNAL, "per-synapse init: the regenerated adjacency does not contain this rank's columns");
		CHECK_CUDA(ctx, cudaMemcpy(c->syn, soa.data(), sizeof(std::uint32_t) * soa.size(), cudaMemcpyHostToDevice));
		c->adj_src = {};
		c->adj_dst = {};
	} else if (c->ops->per_synapse_init) {
		UInt128 const sd = c->init_seed; // the hook's own engine (synapse_population.h:35), drawn at connect() time
		std::vector<long long> off(static_cast<size_t>(src.size) + 1);
		std::vector<std::int32_t> nb(static_cast<size_t>(std::max<long long>(c->edges, 1)));
		CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		CHECK_CUDA(ctx, cudaMemcpy(off.data(), c->offsets, sizeof(long long) * off.size(), cudaMemcpyDeviceToHost));
		if (c->edges)
			CHECK_CUDA(ctx, cudaMemcpy(nb.data(), c->neighbors, sizeof(std::int32_t) * static_cast<size_t>(c->edges), cudaMemcpyDeviceToHost));
		for (auto& d : nb)
			d += static_cast<std::int32_t>(dst.lo);
		std::vector<unsigned char> aos(static_cast<size_t>(std::max<long long>(c->edges, 1)) * c->ops->synapse_bytes);
		c->ops->init_host(c->functor_host.data(), aos.data(), reinterpret_cast<std::int64_t const*>(off.data()), nb.data(), src.size, sd.lo, sd.hi);
		std::vector<std::uint32_t> soa(static_cast<size_t>(words) * static_cast<size_t>(c->syn_stride), 0);
		for (long long e = 0; e < c->edges; e++)
			for (int w = 0; w < words; w++)
				std::memcpy(&soa[static_cast<size_t>(w) * c->syn_stride + e], aos.data() + e * c->ops->synapse_bytes + 4 * w, 4);
		CHECK_CUDA(ctx, cudaMemcpy(c->syn, soa.data(), sizeof(std::uint32_t) * soa.size(), cudaMemcpyHostToDevice));
	} else {
		// one default-constructed synapse, replicated
		std::int64_t const one_off[2] = {0, 1};
		std::int32_t const one_nb  = 0;
		std::vector<unsigned char> one(c->ops->synapse_bytes);
		c->ops->init_host(c->functor_host.data(), one.data(), one_off, &one_nb, 1, 0, 0);
		for (int w = 0; w < words; w++) {
			std::uint32_t v;
			std::memcpy(&v, one.data() + 4 * w, 4);
			fill_u32<<<static_cast<unsigned>((c->syn_stride + 255) / 256), 256, 0, ctx->stream>>>(c->syn + static_cast<size_t>(w) * c->syn_stride, c->syn_stride, v);
			ctx->launches++;
		}
		CHECK_CUDA(ctx, cudaGetLastError());
	}
	return SPICE_OK;
}
} // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

char const* spice_version(void) { return "spice2_b200 0.1 (sm_100a)"; }

int spice_device_check(int device) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || device >= n)
		return SPICE_ERR_NO_DEVICE;
	cudaDeviceProp prop{};
	if (cudaGetDeviceProperties(&prop, device) != cudaSuccess)
		return SPICE_ERR_NO_DEVICE;
	return prop.major == 10 ? SPICE_OK : SPICE_ERR_UNSUPPORTED;
}

char const* spice_last_error(spice_ctx const* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

int spice_ctx_create_seeded(spice_ctx** out, int device, float dt, float max_delay, uint64_t seed_lo, uint64_t seed_hi,
                            int rank, int world, int mode) {
	*out = nullptr;
	if (!(dt > 0) || !(max_delay > 0) || world < 1 || world > kMaxWorld || rank < 0 || rank >= world) {
		g_create_error = "spice_ctx_create: invalid argument";
		return SPICE_ERR_PRECONDITION;
	}
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device >= n) {
		g_create_error = "spice_ctx_create: no CUDA device — this backend has no CPU fallback";
		return SPICE_ERR_NO_DEVICE;
	}
	auto ctx       = std::make_unique<spice_ctx>();
	ctx->device    = device;
	ctx->dt        = dt;
	ctx->max_delay = static_cast<long long>(std::round(max_delay / dt)); // snn.h:18-19
	if (ctx->max_delay < 1) {
		g_create_error = "spice_ctx_create: max_delay must be at least 1 dt";
		return SPICE_ERR_PRECONDITION;
	}
	ctx->seed  = util::seed_seq(UInt128{seed_lo, seed_hi});
	ctx->rank  = rank;
	ctx->world = world;
	ctx->mode  = mode;
	if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
		g_create_error = "spice_ctx_create: cannot create a stream on the device";
		return SPICE_ERR_CUDA;
	}
	ctx->own_stream = true;
	*out            = ctx.release();
	return SPICE_OK;
}

int spice_ctx_create(spice_ctx** out, int device, float dt, float max_delay, uint32_t const* seed_words, int n_seed_words,
                     int rank, int world, int mode) {
	*out = nullptr;
	if (!seed_words || n_seed_words <= 0) {
		g_create_error = "Assertion failed (random.h): il.size() > 0 && \"Please provide at least 1 seed to seed_seq\"";
		return SPICE_ERR_PRECONDITION;
	}
	util::seed_seq const s(seed_words, static_cast<std::size_t>(n_seed_words));
	return spice_ctx_create_seeded(out, device, dt, max_delay, s.seed().lo, s.seed().hi, rank, world, mode);
}

int spice_ctx_destroy(spice_ctx* ctx) {
	if (!ctx)
		return SPICE_OK;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	for (auto& p : ctx->pops) {
		cudaFree(p.functor_dev);
		cudaFree(p.state);
		cudaFree(p.history);
		cudaFree(p.jump_poly);
	}
	for (auto& c : ctx->conns) {
		cudaFree(c.functor_dev);
		cudaFree(c.offsets);
		cudaFree(c.neighbors);
		cudaFree(c.counts);
		cudaFree(c.tile_ptr);
		cudaFree(c.src_snapshot);
		cudaFree(c.packed);
		cudaFree(c.run_ptr);
		cudaFree(c.syn);
		cudaFree(c.ages);
		cudaFree(c.evt_cnt);
		cudaFree(c.evt_off);
		cudaFree(c.evt_fill);
		cudaFree(c.evt_cursor);
		cudaFree(c.evt_list);
	}
	for (int r = 0; r < ctx->world; r++)
		if (r != ctx->rank && ctx->peer_base[r] && ctx->peers_set) {
			// only IPC-opened mappings need closing; same-process peers are plain pointers
			cudaIpcCloseMemHandle(ctx->peer_base[r]);
		}
	cudaFree(ctx->xbase);
	cudaFree(ctx->d_ring_cnt);
	cudaFree(ctx->d_ring_ids);
	for (auto& p : ctx->pops) {
		cudaFree(p.flat_ids);
		cudaFree(p.flat_cnt);
		cudaFreeHost(p.h_stage);
		cudaFreeHost(p.h_stage_cnt);
		for (auto e : p.stage_done)
			if (e)
				cudaEventDestroy(e);
	}
	cudaFree(ctx->d_flat_ids);
	cudaFree(ctx->d_flat_cnt);
	cudaFree(ctx->d_ring_cap);
	cudaFree(ctx->d_seg_lo);
	cudaFree(ctx->d_peer_cnt);
	cudaFree(ctx->d_nib);
	cudaFree(ctx->d_conn_desc);
	cudaFree(ctx->d_work);
	if (ctx->sink_stream) {
		cudaStreamSynchronize(ctx->sink_stream);
		cudaStreamDestroy(ctx->sink_stream);
		cudaEventDestroy(ctx->ev_sink_ready);
		cudaEventDestroy(ctx->ev_sink_done);
	}
	if (ctx->dstream) {
		cudaStreamSynchronize(ctx->dstream);
		cudaStreamDestroy(ctx->dstream);
		cudaEventDestroy(ctx->ev_upd);
		for (auto e : ctx->ev_del)
			cudaEventDestroy(e);
	}
	cudaFree(ctx->d_stats);
	cudaFree(ctx->d_error);
	cudaFree(ctx->d_pop_size);
	cudaFree(ctx->d_sink_cursor);
	cudaFree(ctx->d_sink_stage);
	cudaFreeHost(ctx->h_sink_ids);
	cudaFreeHost(ctx->h_sink_cnt);
	cudaFreeHost(ctx->h_sink_off);
	cudaFreeHost(ctx->h_sink_consumed);
	for (auto& m : ctx->sink_marks)
		cudaEventDestroy(m.second);
	for (auto e : ctx->sink_pool)
		cudaEventDestroy(e);
	for (auto e : ctx->prof_events)
		cudaEventDestroy(e);
	cudaGetLastError();
	for (int k = 0; k < spice_ctx::kAux; k++) {
		if (ctx->aux[k])
			cudaStreamDestroy(ctx->aux[k]);
		if (ctx->ev_join[k])
			cudaEventDestroy(ctx->ev_join[k]);
	}
	if (ctx->ev_fork)
		cudaEventDestroy(ctx->ev_fork);
	if (ctx->own_stream)
		cudaStreamDestroy(ctx->stream);
	delete ctx;
	return SPICE_OK;
}

int spice_ctx_set_stream(spice_ctx* ctx, void* cuda_stream) {
	if (ctx->own_stream) {
		cudaStreamSynchronize(ctx->stream);
		cudaStreamDestroy(ctx->stream);
		ctx->own_stream = false;
	}
	ctx->stream = static_cast<cudaStream_t>(cuda_stream);
	return SPICE_OK;
}

namespace {
// ops table of a host-fed population: stateless, no draws, nothing to launch
int host_pop_noop_update(update_args const*) { return 0; }
int host_pop_noop_export(export_args const*) { return 0; }
int host_pop_noop_import(import_args const*) { return 0; }
void host_pop_noop_init(void const*, void*, std::int64_t, std::uint64_t, std::uint64_t) {}
spice_neuron_ops const host_pop_ops{1, "host", 0, 0, 0, 0, &host_pop_noop_init, &host_pop_noop_update, &host_pop_noop_export, &host_pop_noop_import};
}

int spice_add_host_population(spice_ctx* ctx, int64_t size, spice_host_update_fn update, void* user, int* pop_out) {
	PRE(ctx, update != nullptr);
	PRE(ctx, size >= 0 && size < 2147483647);
	PRE(ctx, !ctx->finalized && "add_population() after the first step is not supported");
	CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	population p;
	p.ops         = &host_pop_ops;
	p.size        = size;
	p.lo          = 0;
	p.hi          = size;
	// More than one rank: a host-fed population is REPLICATED, not partitioned — every rank calls its own copy of the functor
	// (which must emit the same spikes on every rank, as the reference's single instance would) and fills its own ring, as
	// if rank 0 owned every neuron: no spike of this population crosses NVLink.
	p.bounds.assign(static_cast<size_t>(ctx->world) + 1, size);
	p.bounds[0] = 0;
	p.stride      = static_cast<long long>(align_up(static_cast<size_t>(std::max<long long>(size, 1)), 32));
	p.host_update = update;
	p.host_user   = user;
	ctx->pops.push_back(std::move(p));
	if (pop_out)
		*pop_out = static_cast<int>(ctx->pops.size()) - 1;
	return SPICE_OK;
}

int spice_add_population(spice_ctx* ctx, spice_neuron_ops const* ops, int64_t size, void const* functor, int* pop_out) {
	PRE(ctx, ops != nullptr && ops->abi_version == 1);
	PRE(ctx, size >= 0 && size < 2147483647);
	PRE(ctx, !ctx->finalized && "add_population() after the first step is not supported");
	CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	population p;
	p.ops    = ops;
	p.size   = size;
	if (!ctx->next_bounds.empty()) { // spice_set_next_partition: in-degree-balanced (or any other) target ranges
		p.bounds = std::move(ctx->next_bounds); // consumed by this call whether or not it is valid
		ctx->next_bounds.clear();
		PRE(ctx, static_cast<int>(p.bounds.size()) == ctx->world + 1 && p.bounds.front() == 0 && p.bounds.back() == size &&
		             std::is_sorted(p.bounds.begin(), p.bounds.end()));
	} else {
		p.bounds.resize(static_cast<size_t>(ctx->world) + 1);
		for (int r = 0; r <= ctx->world; r++)
			p.bounds[static_cast<size_t>(r)] = size * r / ctx->world;
	}
	p.lo     = p.bounds[static_cast<size_t>(ctx->rank)];
	p.hi     = p.bounds[static_cast<size_t>(ctx->rank) + 1];
	p.stride = static_cast<long long>(align_up(static_cast<size_t>(std::max<long long>(p.hi - p.lo, 1)), 32));
	std::vector<unsigned char> zero(std::max<std::uint32_t>(ops->functor_bytes, 1), 0);
	unsigned char const* f = functor ? static_cast<unsigned char const*>(functor) : zero.data();
	p.functor_host.assign(f, f + ops->functor_bytes);
	CHECK_CUDA(ctx, cudaMalloc(&p.functor_dev, std::max<size_t>(ops->functor_bytes, 16)));
	CHECK_CUDA(ctx, cudaMemcpy(p.functor_dev, f, ops->functor_bytes, cudaMemcpyHostToDevice));
	if (ops->neuron_bytes) {
		// the stateful adapter consumes one seed++ whether or not the model has an init hook
		// (neuron_population.h:60); the hook runs over the WHOLE population with that engine
		UInt128 const sd = (ctx->seed++).seed();
		std::vector<unsigned char> aos(static_cast<size_t>(std::max<long long>(size, 1)) * ops->neuron_bytes);
		ops->init_host(f, aos.data(), size, sd.lo, sd.hi);
		int const words = (ops->neuron_bytes + 3) / 4; // word-SoA; a size that is no multiple of 4 pads its last word
		std::vector<std::uint32_t> soa(static_cast<size_t>(words) * static_cast<size_t>(p.stride), 0);
		for (long long i = p.lo; i < p.hi; i++)
			for (int w = 0; w < words; w++)
				std::memcpy(&soa[static_cast<size_t>(w) * p.stride + (i - p.lo)], aos.data() + i * ops->neuron_bytes + 4 * w,
				            std::min<size_t>(4, ops->neuron_bytes - 4 * w));
		CHECK_CUDA(ctx, cudaMalloc(&p.state, sizeof(std::uint32_t) * soa.size()));
		CHECK_CUDA(ctx, cudaMemcpy(p.state, soa.data(), sizeof(std::uint32_t) * soa.size(), cudaMemcpyHostToDevice));
	}
	ctx->pops.push_back(std::move(p));
	if (pop_out)
		*pop_out = static_cast<int>(ctx->pops.size()) - 1;
	return SPICE_OK;
}

int spice_set_next_partition(spice_ctx* ctx, int64_t const* bounds) {
	ctx->next_bounds.clear();
	if (bounds)
		ctx->next_bounds.assign(bounds, bounds + ctx->world + 1);
	return SPICE_OK;
}

int spice_balance_ranges(int64_t const* weight, int64_t n, int world, int64_t* bounds) {
	if (n < 0 || world < 1 || !bounds || (n > 0 && !weight))
		return SPICE_ERR_PRECONDITION;
	// static synapse-count load balancing: rank r's range ends where the running sum of the in-degrees first reaches
	// (r + 1) / world of the total (prefix sum of the in-degree histogram); every neuron also counts as one
	// unit of update work, which keeps ranges of unconnected neurons from collapsing onto one rank
	long double total = 0;
	for (int64_t i = 0; i < n; i++) {
		if (weight[i] < 0)
			return SPICE_ERR_PRECONDITION;
		total += static_cast<long double>(weight[i]) + 1;
	}
	bounds[0]        = 0;
	long double run  = 0;
	int64_t i        = 0;
	for (int r = 1; r < world; r++) {
		long double const want = total * r / world;
		while (i < n && run + (static_cast<long double>(weight[i]) + 1) / 2 < want)
			run += static_cast<long double>(weight[i++]) + 1;
		bounds[r] = i;
	}
	bounds[world] = n;
	return SPICE_OK;
}

int64_t spice_population_size(spice_ctx const* ctx, int pop) {
	return (pop >= 0 && pop < static_cast<int>(ctx->pops.size())) ? ctx->pops[pop].size : -1;
}

int spice_population_range(spice_ctx const* ctx, int pop, int64_t* lo, int64_t* hi) {
	if (pop < 0 || pop >= static_cast<int>(ctx->pops.size()))
		return SPICE_ERR_PRECONDITION;
	*lo = ctx->pops[pop].lo;
	*hi = ctx->pops[pop].hi;
	return SPICE_OK;
}

int spice_connect_fixed_probability_fast(spice_ctx* ctx, spice_synapse_ops const* ops, int src_pop, int dst_pop, double p, float delay,
                                         void const* functor, int* conn_out) {
	int idx      = -1;
	int const rc = spice_connect_fixed_probability(ctx, ops, src_pop, dst_pop, p, delay, functor, &idx);
	if (rc != SPICE_OK)
		return rc;
	ctx->conns[static_cast<size_t>(idx)].fp_fast = true;
	if (conn_out)
		*conn_out = idx;
	return SPICE_OK;
}

int spice_connect_fixed_probability(spice_ctx* ctx, spice_synapse_ops const* ops, int src_pop, int dst_pop, double p,
                                    float delay, void const* functor, int* conn_out) {
	PRE(ctx, 0 <= p && p <= 1); // topology.cpp:73
	CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	connection c;
	int rc = add_connection_common(ctx, ops, src_pop, dst_pop, delay, functor, &c);
	if (rc != SPICE_OK)
		return rc;
	// synapse_population ctor: _graph(c, seed++) (synapse_population.h:30-31), then the init hook's engine (:35)
	c.fp_seed = (ctx->seed++).seed();
	if (c.stateful && ops->per_synapse_init) {
		c.init_seed = (ctx->seed++).seed();
	}
	c.fp_p       = p;
	c.from_fp    = true;
	c.pending_fp = true; // generated by finalize(), together with the network's other connections
	ctx->conns.push_back(std::move(c));
	if (conn_out)
		*conn_out = static_cast<int>(ctx->conns.size()) - 1;
	return SPICE_OK;
}

int spice_connect_adj_list(spice_ctx* ctx, spice_synapse_ops const* ops, int src_pop, int dst_pop, int32_t const* edges_src,
                           int32_t const* edges_dst, int64_t n_edges, float delay, void const* functor, int* conn_out) {
	CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	connection c;
	int rc = add_connection_common(ctx, ops, src_pop, dst_pop, delay, functor, &c);
	if (rc != SPICE_OK)
		return rc;
	(void)(ctx->seed++); // the graph still consumes its seed (synapse_population.h:31)
	population const& src = ctx->pops[src_pop];
	population const& dst = ctx->pops[dst_pop];
	// adj_list::generate: sort packed (src << 32 | dst) and stream into CSR (topology.cpp:63-71) — on the device (radix
	// sort of the packed keys, rows by histogram + scan: generator.cu generate_adj_list)
	{
		gen::result r;
		std::string err;
		bool dup = false;
		CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		int const grc = gen::generate_adj_list(ctx->stream, edges_src, edges_dst, n_edges, src.size, dst.size, dst.lo, dst.hi, &r, &dup, &err);
		if (grc != 0) {
			cudaFree(r.offsets);
			cudaFree(r.neighbors);
			drop_connection(c);
			if (grc == 1) // the reference's SPICE_PRE on every streamed edge (topology.cpp:16-18)
				return fail(ctx, SPICE_ERR_PRECONDITION, std::string("Assertion failed (") + __FILE__ + ":" + std::to_string(__LINE__) +
				                                            "): 0 <= src && src < src_count && 0 <= dst && dst < dst_count");
			return fail(ctx, grc, err);
		}
		c.offsets    = r.offsets;
		c.neighbors  = r.neighbors;
		c.edges      = r.edges;
		c.duplicates = dup; // a multapse: rows may repeat a target
		ctx->launches += r.launches;
	}
	if (c.stateful && ops->per_synapse_init) {
		c.init_seed = (ctx->seed++).seed();
		if (ctx->world > 1) { // the hook runs over the whole adjacency (init_synapses)
			c.adj_src.assign(edges_src, edges_src + n_edges);
			c.adj_dst.assign(edges_dst, edges_dst + n_edges);
		}
	}
	rc = init_synapses(ctx, &c);
	if (rc != SPICE_OK) {
		drop_connection(c);
		return rc;
	}
	ctx->conns.push_back(std::move(c));
	if (conn_out)
		*conn_out = static_cast<int>(ctx->conns.size()) - 1;
	return SPICE_OK;
}

int spice_connection_csr(spice_ctx* ctx, int conn, int64_t* n_edges_out, int64_t* offsets_out, int32_t* neighbors_out) {
	PRE(ctx, conn >= 0 && conn < static_cast<int>(ctx->conns.size()));
	CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	{
		int const rc = finalize(ctx); // the adjacency exists once the network has been built
		if (rc != SPICE_OK)
			return rc;
	}
	connection const& c = ctx->conns[conn];
	if (n_edges_out)
		*n_edges_out = c.edges;
	CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	if (offsets_out)
		CHECK_CUDA(ctx, cudaMemcpy(offsets_out, c.offsets, sizeof(long long) * static_cast<size_t>(ctx->pops[c.src].size + 1), cudaMemcpyDeviceToHost));
	if (neighbors_out && c.edges) {
		std::int32_t const* from = c.neighbors;
		std::int32_t* tmp        = nullptr;
		if (c.arranged) { // the delivery kernel's stream -> ascending local columns
			CHECK_CUDA(ctx, cudaMalloc(&tmp, sizeof(std::int32_t) * static_cast<size_t>(c.edges)));
			int const e = deliver::unpack_rows(ctx->stream, c.packed, c.run_ptr, c.offsets, ctx->pops[c.src].size, c.tile, c.tiles, ctx->tile_cap, tmp);
			if (e != 0) {
				cudaFree(tmp);
				return fail(ctx, SPICE_ERR_CUDA, std::string("unpack_rows: ") + cudaGetErrorString(static_cast<cudaError_t>(e)));
			}
			from = tmp;
		}
		cudaError_t const e = cudaMemcpyAsync(neighbors_out, from, sizeof(std::int32_t) * static_cast<size_t>(c.edges), cudaMemcpyDeviceToHost, ctx->stream);
		cudaError_t const e2 = cudaStreamSynchronize(ctx->stream);
		cudaFree(tmp);
		CHECK_CUDA(ctx, e);
		CHECK_CUDA(ctx, e2);
	}
	return SPICE_OK;
}

int spice_connection_synapses(spice_ctx* ctx, int conn, void* out, int64_t bytes) {
	PRE(ctx, conn >= 0 && conn < static_cast<int>(ctx->conns.size()));
	{
		int const rc = finalize(ctx);
		if (rc != SPICE_OK)
			return rc;
	}
	connection const& c = ctx->conns[conn];
	if (!c.stateful)
		return fail(ctx, SPICE_ERR_UNSUPPORTED, "stateless connection: no per-synapse state");
	PRE(ctx, bytes == c.edges * static_cast<int64_t>(c.ops->synapse_bytes));
	CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	int const words = static_cast<int>(c.ops->synapse_bytes / 4);
	std::vector<std::uint32_t> soa(static_cast<size_t>(words) * static_cast<size_t>(c.syn_stride));
	CHECK_CUDA(ctx, cudaMemcpy(soa.data(), c.syn, sizeof(std::uint32_t) * soa.size(), cudaMemcpyDeviceToHost));
	auto* o = static_cast<unsigned char*>(out);
	for (long long e = 0; e < c.edges; e++)
		for (int w = 0; w < words; w++)
			std::memcpy(o + e * c.ops->synapse_bytes + 4 * w, &soa[static_cast<size_t>(w) * c.syn_stride + e], 4);
	return SPICE_OK;
}

int spice_ctx_finalize(spice_ctx* ctx) { return finalize(ctx); }

int spice_ctx_peer_handle(spice_ctx* ctx, void* out, int64_t* bytes) {
	int rc = finalize(ctx);
	if (rc != SPICE_OK)
		return rc;
	if (bytes)
		*bytes = sizeof(peer_blob);
	if (!out)
		return SPICE_OK;
	peer_blob b{};
	b.magic   = 0x5350494345423230ull;
	b.pid     = static_cast<int>(getpid());
	b.device  = ctx->device;
	b.rank    = ctx->rank;
	b.bytes   = ctx->xbytes;
	b.raw_ptr = reinterpret_cast<unsigned long long>(ctx->xbase);
	CHECK_CUDA(ctx, cudaIpcGetMemHandle(&b.handle, ctx->xbase));
	std::memcpy(out, &b, sizeof b);
	return SPICE_OK;
}

int spice_ctx_set_peers(spice_ctx* ctx, void const* blobs, int64_t bytes_each) {
	PRE(ctx, ctx->finalized);
	PRE(ctx, bytes_each == static_cast<int64_t>(sizeof(peer_blob)));
	CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	for (int r = 0; r < ctx->world; r++) {
		peer_blob b;
		std::memcpy(&b, static_cast<unsigned char const*>(blobs) + r * sizeof(peer_blob), sizeof b);
		PRE(ctx, b.magic == 0x5350494345423230ull && b.rank == r && b.bytes == ctx->xbytes);
		if (r == ctx->rank)
			continue;
		if (b.pid == static_cast<int>(getpid())) {
			// same process (tests, single-process multi-device): the raw pointer is directly usable
			if (b.device != ctx->device) {
				cudaError_t const e = cudaDeviceEnablePeerAccess(b.device, 0);
				if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
					CHECK_CUDA(ctx, e);
				cudaGetLastError();
			}
			ctx->peer_base[r] = reinterpret_cast<unsigned char*>(b.raw_ptr);
		} else {
			void* p = nullptr;
			CHECK_CUDA(ctx, cudaIpcOpenMemHandle(&p, b.handle, cudaIpcMemLazyEnablePeerAccess));
			ctx->peer_base[r] = static_cast<unsigned char*>(p);
			ctx->peers_set    = true;
		}
	}
	return upload_peer_tables(ctx);
}

int spice_run(spice_ctx* ctx, int64_t n_steps) {
	PRE(ctx, n_steps >= 0);
	CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	int rc = finalize(ctx);
	if (rc != SPICE_OK)
		return rc;
	if (ctx->world > 1)
		for (int r = 0; r < ctx->world; r++)
			PRE(ctx, ctx->peer_base[r] != nullptr && "multi-rank context: call spice_ctx_set_peers() first");
	while (n_steps > 0) {
		int const n = static_cast<int>(std::min<int64_t>(n_steps, ctx->window));
		rc          = run_window(ctx, n);
		if (rc != SPICE_OK)
			return rc;
		n_steps -= n;
	}
	if (ctx->pipelined) // whatever follows on the context's stream (readouts, the next run) sees every delivery of this one
		for (int p = 0; p < 2; p++)
			if (ctx->del_pending[p])
				CHECK_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_del[p], 0));
	return SPICE_OK;
}

int spice_sync(spice_ctx* ctx) {
	CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	if (!ctx->finalized) {
		CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		return SPICE_OK;
	}
	return check_device_error(ctx);
}

int64_t spice_time(spice_ctx const* ctx) { return ctx->time; }

int spice_spikes(spice_ctx* ctx, int pop, int64_t age, int32_t const** ids_out, int64_t* n_out) {
	PRE(ctx, pop >= 0 && pop < static_cast<int>(ctx->pops.size()));
	// neuron_population.h:148: 0 <= age < number of steps held (at most max_delay)
	PRE(ctx, 0 <= age && age < std::min<long long>(ctx->time, ctx->max_delay));
	CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	int rc = check_device_error(ctx);
	if (rc != SPICE_OK)
		return rc;
	population const& p  = ctx->pops[pop];
	long long const step = ctx->time - 1 - age;
	long long const slot = step % ctx->ring;
	host_spikes& hs      = ctx->spike_cache[pop][static_cast<size_t>(slot)];
	if (hs.step != step) {
		std::vector<std::uint32_t> cnt(ctx->world);
		CHECK_CUDA(ctx, cudaMemcpy(cnt.data(), xptr<std::uint32_t>(ctx->xbase, p.ring_cnt_off) + slot * ctx->world,
		                           sizeof(std::uint32_t) * ctx->world, cudaMemcpyDeviceToHost));
		long long total = 0;
		for (auto c : cnt)
			total += c;
		hs.ids.resize(static_cast<size_t>(total));
		long long at = 0;
		for (int r = 0; r < ctx->world; r++) {
			if (cnt[r]) {
				CHECK_CUDA(ctx, cudaMemcpy(hs.ids.data() + at,
				                           xptr<std::int32_t>(ctx->xbase, p.ring_ids_off) + slot * std::max<long long>(p.size, 1) + p.seg_lo[r],
				                           sizeof(std::int32_t) * cnt[r], cudaMemcpyDeviceToHost));
				// a segment holds one rank's spikes in arrival order; the reference lists them ascending
				std::sort(hs.ids.begin() + at, hs.ids.begin() + at + cnt[r]);
			}
			at += cnt[r];
		}
		hs.step = step;
	}
	*ids_out = hs.ids.data();
	*n_out   = static_cast<int64_t>(hs.ids.size());
	return SPICE_OK;
}

int spice_neurons(spice_ctx* ctx, int pop, void* out, int64_t bytes) {
	PRE(ctx, pop >= 0 && pop < static_cast<int>(ctx->pops.size()));
	population& p = ctx->pops[pop];
	PRE(ctx, p.ops->neuron_bytes != 0 && "Can only return collections of stateful neurons.");
	long long const n = p.hi - p.lo;
	PRE(ctx, bytes == n * p.ops->neuron_bytes);
	CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	int rc = finalize(ctx);
	if (rc != SPICE_OK)
		return rc;
	if (n == 0)
		return SPICE_OK;
	void* dev = nullptr;
	CHECK_CUDA(ctx, cudaMalloc(&dev, static_cast<size_t>(bytes)));
	export_args ea{};
	ea.stream  = ctx->stream;
	ea.state   = p.state;
	ea.n_local = n;
	ea.stride  = p.stride;
	ea.t_next  = ctx->time;
	fill_incoming(ctx, p, ea.in, &ea.n_in);
	ea.out_aos = dev;
	int const e = p.ops->launch_export(&ea);
	ctx->launches++;
	if (e != 0) {
		cudaFree(dev);
		return fail(ctx, SPICE_ERR_CUDA, std::string("export launch: ") + cudaGetErrorString(static_cast<cudaError_t>(e)));
	}
	cudaError_t ce = cudaMemcpyAsync(out, dev, static_cast<size_t>(bytes), cudaMemcpyDeviceToHost, ctx->stream);
	if (ce == cudaSuccess)
		ce = cudaStreamSynchronize(ctx->stream);
	cudaFree(dev);
	CHECK_CUDA(ctx, ce);
	return SPICE_OK;
}

int spice_set_neurons(spice_ctx* ctx, int pop, void const* in, int64_t bytes) {
	PRE(ctx, pop >= 0 && pop < static_cast<int>(ctx->pops.size()));
	population& p = ctx->pops[pop];
	PRE(ctx, p.ops->neuron_bytes != 0);
	long long const n = p.hi - p.lo;
	PRE(ctx, bytes == n * p.ops->neuron_bytes);
	CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	if (n == 0)
		return SPICE_OK;
	void* dev = nullptr;
	CHECK_CUDA(ctx, cudaMalloc(&dev, static_cast<size_t>(bytes)));
	cudaError_t ce = cudaMemcpyAsync(dev, in, static_cast<size_t>(bytes), cudaMemcpyHostToDevice, ctx->stream);
	import_args ia{};
	ia.stream  = ctx->stream;
	ia.state   = p.state;
	ia.n_local = n;
	ia.stride  = p.stride;
	ia.in_aos  = dev;
	ia.t_next  = ctx->time;
	ia.n_in    = 0;
	if (ctx->finalized)
		fill_incoming(ctx, p, ia.in, &ia.n_in);
	if (ce == cudaSuccess)
		ce = static_cast<cudaError_t>(p.ops->launch_import(&ia));
	ctx->launches++;
	if (ce == cudaSuccess)
		ce = cudaStreamSynchronize(ctx->stream);
	cudaFree(dev);
	CHECK_CUDA(ctx, ce);
	return SPICE_OK;
}

int spice_raster_enable(spice_ctx* ctx, int enable) {
	CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	if (enable && !ctx->finalized) {
		int const rc = finalize(ctx);
		if (rc != SPICE_OK)
			return rc;
	}
	if (enable && !ctx->d_sink_cursor) {
		long long total = 0, largest = 1;
		for (auto const& p : ctx->pops) {
			total += p.size;
			largest = std::max(largest, p.size);
		}
		int const np        = static_cast<int>(std::max<size_t>(ctx->pops.size(), 1));
		ctx->sink_steps_cap = 1 << 15;
		ctx->sink_cap       = std::max<long long>(1 << 22, std::min<long long>(total * 16, 1ll << 25));
		if (char const* e = std::getenv("SPICE_SINK_IDS")) // ring sizes, for tests of the wrap-around
			ctx->sink_cap = std::max<long long>(1024, std::atoll(e));
		if (char const* e = std::getenv("SPICE_SINK_STEPS"))
			ctx->sink_steps_cap = std::max<long long>(ctx->window, std::atoll(e));
		long long const W   = std::min<long long>(31, ((largest + 32 * kSinkThreads - 1) / (32 * kSinkThreads)) | 1);
		ctx->sink_smem      = static_cast<int>(W * kSinkThreads * 4);
		CHECK_CUDA(ctx, cudaFuncSetAttribute(sink_pack, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->sink_smem));
		CHECK_CUDA(ctx, cudaMalloc(&ctx->d_sink_cursor, 2 * sizeof(unsigned long long)));
		CHECK_CUDA(ctx, cudaMemset(ctx->d_sink_cursor, 0, 2 * sizeof(unsigned long long)));
		CHECK_CUDA(ctx, cudaMalloc(&ctx->d_sink_stage, sizeof(std::int32_t) * static_cast<size_t>(ctx->sink_cap)));
		CHECK_CUDA(ctx, cudaHostAlloc(&ctx->h_sink_ids, sizeof(std::int32_t) * static_cast<size_t>(ctx->sink_cap), cudaHostAllocMapped));
		CHECK_CUDA(ctx, cudaHostAlloc(&ctx->h_sink_cnt, sizeof(long long) * static_cast<size_t>(ctx->sink_steps_cap) * np, cudaHostAllocMapped));
		CHECK_CUDA(ctx, cudaHostAlloc(&ctx->h_sink_off, sizeof(unsigned long long) * static_cast<size_t>(ctx->sink_steps_cap) * np, cudaHostAllocMapped));
		CHECK_CUDA(ctx, cudaHostAlloc(&ctx->h_sink_consumed, 4 * sizeof(unsigned long long), cudaHostAllocMapped));
		std::memset(ctx->h_sink_consumed, 0, 4 * sizeof(unsigned long long));
		ctx->h_sink_error = reinterpret_cast<int*>(ctx->h_sink_consumed + 2);
	}
	ctx->raster_on = enable != 0;
	return SPICE_OK;
}

namespace {
// block until the first `upto` logged steps are in host memory
int sink_wait(spice_ctx* ctx, long long upto) {
	if (upto > ctx->sink_steps_complete) {
		size_t k = 0; // the first mark that covers `upto` (marks are ascending)
		while (k < ctx->sink_marks.size() && ctx->sink_marks[k].first < upto)
			k++;
		if (k < ctx->sink_marks.size()) {
			CHECK_CUDA(ctx, cudaEventSynchronize(ctx->sink_marks[k].second));
			ctx->sink_steps_complete = ctx->sink_marks[k].first;
			for (size_t i = 0; i <= k; i++) {
				ctx->sink_pool.push_back(ctx->sink_marks.front().second);
				ctx->sink_marks.pop_front();
			}
		} else { // its mark was recycled
			CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
			if (ctx->sink_stream)
				CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->sink_stream));
			ctx->sink_steps_complete = ctx->sink_steps_issued;
		}
	}
	int const h = ctx->h_sink_error ? *reinterpret_cast<int volatile*>(ctx->h_sink_error) : 0;
	if (h & 4)
		return fail(ctx, SPICE_ERR_INTERNAL, "raster log: step capacity exceeded (read the raster more often)");
	if (h & 8)
		return fail(ctx, SPICE_ERR_INTERNAL, "raster log: id capacity exceeded (read the raster more often)");
	return SPICE_OK;
}
}

int spice_raster_size(spice_ctx* ctx, int64_t max_steps, int64_t* n_steps_out, int64_t* n_ids_out) {
	CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	long long const avail = ctx->sink_steps_issued - ctx->sink_steps_read;
	long long const n     = max_steps > 0 ? std::min<long long>(max_steps, avail) : avail;
	long long ids         = 0;
	if (n > 0) {
		int const rc = sink_wait(ctx, ctx->sink_steps_read + n);
		if (rc != SPICE_OK)
			return rc;
		int const np = static_cast<int>(ctx->pops.size());
		for (long long i = 0; i < n; i++)
			for (int p = 0; p < np; p++)
				ids += ctx->h_sink_cnt[((ctx->sink_steps_read + i) % ctx->sink_steps_cap) * np + p];
	}
	if (n_steps_out)
		*n_steps_out = n;
	if (n_ids_out)
		*n_ids_out = ids;
	return SPICE_OK;
}

int spice_raster_read(spice_ctx* ctx, int64_t n_steps, int64_t* counts_out, int32_t* ids_out) {
	CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	int64_t steps = 0, nids = 0;
	int const rc = spice_raster_size(ctx, n_steps, &steps, &nids);
	if (rc != SPICE_OK)
		return rc;
	PRE(ctx, n_steps <= 0 || steps == n_steps);
	int const np = static_cast<int>(ctx->pops.size());
	if (steps == 0 || np == 0)
		return SPICE_OK;
	for (long long i = 0; i < steps; i++)
		for (int p = 0; p < np; p++)
			counts_out[i * np + p] = ctx->h_sink_cnt[((ctx->sink_steps_read + i) % ctx->sink_steps_cap) * np + p];
	// the lists are consecutive in (step, population) order: one range of the ring, at most two pieces
	unsigned long long const first = ctx->h_sink_off[(ctx->sink_steps_read % ctx->sink_steps_cap) * np];
	unsigned long long const cap   = static_cast<unsigned long long>(ctx->sink_cap);
	unsigned long long const at    = first % cap;
	unsigned long long const head  = std::min<unsigned long long>(static_cast<unsigned long long>(nids), cap - at);
	std::memcpy(ids_out, ctx->h_sink_ids + at, sizeof(std::int32_t) * head);
	std::memcpy(ids_out + head, ctx->h_sink_ids, sizeof(std::int32_t) * (static_cast<unsigned long long>(nids) - head));
	ctx->sink_steps_read += steps;
	ctx->sink_ids_read = first + static_cast<unsigned long long>(nids);
	reinterpret_cast<unsigned long long volatile*>(ctx->h_sink_consumed)[0] = static_cast<unsigned long long>(ctx->sink_steps_read);
	reinterpret_cast<unsigned long long volatile*>(ctx->h_sink_consumed)[1] = ctx->sink_ids_read;
	return SPICE_OK;
}

int spice_stats(spice_ctx* ctx, int64_t* synaptic_events, int64_t* spikes_delivered, int64_t* kernel_launches) {
	CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	unsigned long long h[2] = {0, 0};
	if (ctx->d_stats) {
		CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		CHECK_CUDA(ctx, cudaMemcpy(h, ctx->d_stats, sizeof h, cudaMemcpyDeviceToHost));
	}
	if (synaptic_events)
		*synaptic_events = static_cast<int64_t>(h[0]);
	if (spikes_delivered)
		*spikes_delivered = static_cast<int64_t>(h[1]);
	if (kernel_launches)
		*kernel_launches = ctx->launches;
	return SPICE_OK;
}

int64_t spice_windows_run(spice_ctx const* ctx) { return ctx->windows_run; }

int spice_profile_enable(spice_ctx* ctx, int enable) {
	ctx->profile       = enable != 0;
	ctx->profile_every = enable > 1 ? enable : 1;
	return SPICE_OK;
}

int spice_profile_read(spice_ctx* ctx, double* update_ms, double* deliver_ms, double* exchange_ms, int64_t* windows) {
	CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	int const rc = prof_collect(ctx);
	if (rc != SPICE_OK)
		return rc;
	if (update_ms)
		*update_ms = ctx->prof_update;
	if (deliver_ms)
		*deliver_ms = ctx->prof_deliver;
	if (exchange_ms)
		*exchange_ms = ctx->prof_exchange;
	if (windows)
		*windows = ctx->prof_windows;
	ctx->prof_update = ctx->prof_deliver = ctx->prof_exchange = 0;
	ctx->prof_windows = 0;
	return SPICE_OK;
}

// ---- standalone generation ---------------------------------------------------------------------
struct spice_adjacency {
	int device;
	long long src;
	gen::result r;
};

int64_t spice_fixed_probability_max_degree(int64_t dst_count, double p) { return gen::max_degree(dst_count, p); }

int spice_fixed_probability_generate(int device, int64_t src_count, int64_t dst_count, double p, uint64_t seed_lo,
                                     uint64_t seed_hi, int64_t col_lo, int64_t col_hi, spice_adjacency** out) {
	*out = nullptr;
	if (!(0 <= p && p <= 1) || src_count < 0 || dst_count < 0 || src_count >= 2147483647 || dst_count >= 2147483647 ||
	    col_lo < 0 || col_hi > dst_count || col_lo > col_hi) {
		g_create_error = "spice_fixed_probability_generate: invalid argument";
		return SPICE_ERR_PRECONDITION;
	}
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device >= n) {
		g_create_error = "no CUDA device — this backend has no CPU fallback";
		return SPICE_ERR_NO_DEVICE;
	}
	cudaSetDevice(device);
	auto a    = std::make_unique<spice_adjacency>();
	a->device = device;
	a->src    = src_count;
	std::string err;
	int const rc = gen::generate_fixed_probability(nullptr, src_count, dst_count, p, seed_lo, seed_hi, col_lo, col_hi, 0, &a->r, &err);
	if (rc != 0) {
		cudaFree(a->r.offsets);
		cudaFree(a->r.neighbors);
		g_create_error = err;
		return rc;
	}
	*out = a.release();
	return SPICE_OK;
}

int spice_fixed_probability_generate_fast(int device, int64_t src_count, int64_t dst_count, double p, uint64_t seed_lo, uint64_t seed_hi,
                                          int64_t col_lo, int64_t col_hi, spice_adjacency** out) {
	*out = nullptr;
	if (!(0 <= p && p <= 1) || src_count < 0 || dst_count < 0 || src_count >= 2147483647 || dst_count >= 2147483647 || col_lo < 0 ||
	    col_hi > dst_count || col_lo > col_hi) {
		g_create_error = "spice_fixed_probability_generate_fast: invalid argument";
		return SPICE_ERR_PRECONDITION;
	}
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device >= n) {
		g_create_error = "no CUDA device — this backend has no CPU fallback";
		return SPICE_ERR_NO_DEVICE;
	}
	cudaSetDevice(device);
	auto a    = std::make_unique<spice_adjacency>();
	a->device = device;
	a->src    = src_count;
	std::string err;
	int const rc = gen::generate_fixed_probability_fast(nullptr, src_count, dst_count, p, seed_lo, seed_hi, col_lo, col_hi, &a->r, &err);
	if (rc != 0) {
		cudaFree(a->r.offsets);
		cudaFree(a->r.neighbors);
		g_create_error = err;
		return rc;
	}
	*out = a.release();
	return SPICE_OK;
}

int spice_adj_list_generate(int device, int32_t const* edges_src, int32_t const* edges_dst, int64_t n_edges, int64_t src_count, int64_t dst_count,
                            int64_t col_lo, int64_t col_hi, spice_adjacency** out) {
	*out = nullptr;
	if (n_edges < 0 || src_count < 0 || dst_count < 0 || src_count >= 2147483647 || dst_count >= 2147483647 || col_lo < 0 ||
	    col_hi > dst_count || col_lo > col_hi || (n_edges > 0 && (!edges_src || !edges_dst))) {
		g_create_error = "spice_adj_list_generate: invalid argument";
		return SPICE_ERR_PRECONDITION;
	}
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device >= n) {
		g_create_error = "no CUDA device — this backend has no CPU fallback";
		return SPICE_ERR_NO_DEVICE;
	}
	cudaSetDevice(device);
	auto a    = std::make_unique<spice_adjacency>();
	a->device = device;
	a->src    = src_count;
	std::string err;
	int const rc = gen::generate_adj_list(nullptr, edges_src, edges_dst, n_edges, src_count, dst_count, col_lo, col_hi, &a->r, nullptr, &err);
	if (rc != 0) {
		cudaFree(a->r.offsets);
		cudaFree(a->r.neighbors);
		g_create_error = rc == 1 ? "Assertion failed (adj_list): 0 <= src && src < src_count && 0 <= dst && dst < dst_count" : err;
		return rc == 1 ? SPICE_ERR_PRECONDITION : rc;
	}
	*out = a.release();
	return SPICE_OK;
}

int64_t spice_adjacency_edges(spice_adjacency const* a) { return a->r.edges; }
void* spice_adjacency_offsets_dev(spice_adjacency const* a) { return a->r.offsets; }
void* spice_adjacency_neighbors_dev(spice_adjacency const* a) { return a->r.neighbors; }

int spice_adjacency_copy(spice_adjacency const* a, int64_t* offsets_host, int32_t* neighbors_host) {
	cudaSetDevice(a->device);
	if (offsets_host && cudaMemcpy(offsets_host, a->r.offsets, sizeof(long long) * static_cast<size_t>(a->src + 1), cudaMemcpyDeviceToHost) != cudaSuccess)
		return SPICE_ERR_CUDA;
	if (neighbors_host && a->r.edges &&
	    cudaMemcpy(neighbors_host, a->r.neighbors, sizeof(std::int32_t) * static_cast<size_t>(a->r.edges), cudaMemcpyDeviceToHost) != cudaSuccess)
		return SPICE_ERR_CUDA;
	return SPICE_OK;
}

int spice_adjacency_copy_range(spice_adjacency const* a, int64_t edge_lo, int64_t edge_hi, int32_t* neighbors_host) {
	if (!a || edge_lo < 0 || edge_hi < edge_lo || edge_hi > a->r.edges || (!neighbors_host && edge_hi > edge_lo))
		return SPICE_ERR_PRECONDITION;
	cudaSetDevice(a->device);
	if (edge_hi > edge_lo && cudaMemcpy(neighbors_host, a->r.neighbors + edge_lo, sizeof(std::int32_t) * static_cast<size_t>(edge_hi - edge_lo),
	                                   cudaMemcpyDeviceToHost) != cudaSuccess)
		return SPICE_ERR_CUDA;
	return SPICE_OK;
}

int spice_adjacency_timing(spice_adjacency const* a, float* total_ms, float* rows_kernel_ms, int64_t* draws) {
	if (total_ms)
		*total_ms = a->r.total_ms;
	if (rows_kernel_ms)
		*rows_kernel_ms = a->r.rows_ms;
	if (draws)
		*draws = a->r.draws;
	return SPICE_OK;
}

int spice_adjacency_destroy(spice_adjacency* a) {
	if (!a)
		return SPICE_OK;
	cudaSetDevice(a->device);
	cudaFree(a->r.offsets);
	cudaFree(a->r.neighbors);
	delete a;
	return SPICE_OK;
}

int spice_ctx_seed(spice_ctx const* ctx, uint64_t out[2]) {
	if (!ctx || !out)
		return SPICE_ERR_PRECONDITION;
	out[0] = ctx->seed.seed().lo;
	out[1] = ctx->seed.seed().hi;
	return SPICE_OK;
}

uint64_t spice_fnv1a64(void const* data, int64_t bytes) {
	auto const* b = static_cast<unsigned char const*>(data);
	uint64_t h    = 0xcbf29ce484222325ull;
	for (int64_t i = 0; i < bytes; i++)
		h = (h ^ b[i]) * 0x100000001b3ull;
	return h;
}

void spice_seed_seq(uint32_t const* words, int n, uint64_t out[2]) {
	util::seed_seq s(words, static_cast<std::size_t>(n));
	out[0] = s.seed().lo;
	out[1] = s.seed().hi;
}

void spice_seed_next(uint64_t seed[2]) {
	util::seed_seq s(UInt128{seed[0], seed[1]});
	s++;
	seed[0] = s.seed().lo;
	seed[1] = s.seed().hi;
}
}
