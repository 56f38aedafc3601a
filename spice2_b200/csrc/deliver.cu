// Spike delivery, tiled: the reference's hot loop B (spice/include/spice/detail/synapse_population.h:
// 88,99,118-133 under spice/src/snn.cpp:21-25) — for src in spikes, for dst in row(src): deliver —
// for all connections whose events are integer counts (every stateless synapse), one launch per
// window of steps.
//
// Why not one atomic per event: a B200 SM retires global reductions at ~1.3 cycles per lane and
// shared-memory atomics at ~2 cycles per lane, i.e. <= 2e11 events/s for the whole GPU, an
// eighth of what HBM can stream (4 B of CSR per event).  The rows of a connection are strictly
// ascending target lists, so the entries of ONE row never collide with each other: a warp that
// works on one row at a time can bump its counters with plain shared-memory load/add/store.
//
// Layout of the work:
//   * the targets of a connection are cut into tiles of <= 5120 neurons; tile_ptr[src][k] says
//     where tile k's share of row src starts (built once per connection), so a tile's share of
//     a row is one contiguous run of ~p * tile entries;
//   * a unit = (connection, step of the window, tile).  A CTA of 4 warps owns a unit.  Every
//     warp keeps its own counters for the tile in 10 KB of shared memory and takes every 4th
//     batch of 32 spikes of the step's spike list; it streams each spiking source's run with
//     one 16-byte load per lane (16 runs in flight, in registers) and counts with non-atomic
//     shared-memory read-modify-writes — inside a warp no barrier and no atomic;
//   * bank conflicts.  A counting instruction scatters 32 lanes over the tile; at random that
//     costs ~3.1 shared-memory wavefronts per instruction instead of 1 and made the first
//     version of this kernel LSU-bound at 39 % of the HBM roofline.  So every target has TWO
//     u8 counters (arrays A and B, whose bank assignments differ by a per-row rotation), and
//     once per connection pack_runs() rewrites each run into the kernel's own stream format:
//     every entry becomes the byte address of one of its target's two counters, chosen
//     (2-choice balancing) and ordered so that the lanes of one instruction fall into
//     different banks; a run is padded to whole 16-byte groups with addresses of a 128-byte
//     dump area (in banks the instruction does not use), so the counting code needs no length,
//     no alignment and no per-entry predicate: a lane either holds a whole group or nothing.
//     The canonical ascending CSR row is recovered by decoding and sorting (unpack_rows);
//   * a u8 counter holds 255: a warp counts at most 224 runs (7 batches) per round; after each
//     round the CTA adds its 4 x 2 arrays and stores (first round) or adds (later rounds) the
//     tile's counters to counts[slot(step + delay)][tile] with plain vector accesses: it is the
//     only writer of that range, and the target's update kernel (the only reader) runs in a
//     later window.  Units are handed out by a global counter to a persistent grid, heaviest
//     connections first;
//   * windows with few, long units (a rank of a multi-GPU run) are handed out round by round
//     instead (deliver_plan.h, plan_items, deliver_tiles<true>): round 0 of a unit stores and
//     publishes a flag, the later rounds wait for it and add with global reductions.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>

#include <cub/block/block_scan.cuh>
#include <cub/device/device_scan.cuh>

#include "deliver.h"
#include "deliver_plan.h"

namespace spice::deliver {
namespace {

constexpr int kWarps       = 4;  // warps per CTA: they share a unit, each with its own copy of the tile
constexpr int kCtasPerSm   = 5;  // register budget: 20 warps per SM
constexpr int kRing        = 16; // runs in flight per warp (16 bytes per lane and run, in registers)
constexpr int kRoundBatches = 7; // batches of 32 runs a warp counts between two merges (224 <= 255: u8 counters)
constexpr unsigned kFull   = 0xffffffffu;

// count the 4 entries of `v` (counter addresses of 4 distinct targets, so the loads may all
// precede the stores); a lane without a group holds v.x < 0
__device__ __forceinline__ void tally(unsigned char* cnt, int4 v) {
	if (v.x >= 0) {
		unsigned char const c0 = cnt[v.x], c1 = cnt[v.y], c2 = cnt[v.z], c3 = cnt[v.w];
		cnt[v.x] = c0 + 1;
		cnt[v.y] = c1 + 1;
		cnt[v.z] = c2 + 1;
		cnt[v.w] = c3 + 1;
	}
}

// zero counters [0, bytes) of array A and of array B (bytes a multiple of 128)
__device__ __forceinline__ void zero_tile(unsigned char* cnt, int cap, int bytes, int lane) {
	uint4* a4 = reinterpret_cast<uint4*>(cnt);
	uint4* b4 = reinterpret_cast<uint4*>(cnt + cap);
	for (int i = lane; i < bytes / 16; i += 32) {
		a4[i] = make_uint4(0, 0, 0, 0);
		b4[i] = make_uint4(0, 0, 0, 0);
	}
	__syncwarp();
}

__device__ __forceinline__ int4 ldg_stream(void const* p) {
	int4 v;
	asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
	return v;
}

// groups [32, ng) of a long run; out of line: the pipeline's loop is unrolled 32 times and should stay small
__device__ __noinline__ void tally_rest(unsigned char* cnt, int4 const* g, unsigned ng, int lane) {
	__syncwarp();
	for (unsigned off = 32; off < ng; off += 32) {
		int4 w = make_int4(-1, 0, 0, 0);
		if (off + lane < ng)
			w = ldg_stream(g + off);
		tally(cnt, w);
	}
	__syncwarp();
}

// One run (a tile's share of one spiking source's row) as the pipeline sees it: groups
// [g0, g0 + ng) of the connection's packed stream (a group = 16 bytes = 4 entries).
struct alignas(8) run_desc {
	unsigned g0, ng;
};

// shared memory of one warp: [descriptors: 2 batches x 32 x 8 B][counters: 2 arrays x tile_cap x 1 B][dump: 128 B]
constexpr int kDescBytes = 2 * 32 * static_cast<int>(sizeof(run_desc));
constexpr int kDumpBytes = 128;
__host__ __device__ constexpr size_t warp_smem(int tile_cap) {
	return static_cast<size_t>(kDescBytes) + 2 * static_cast<size_t>(tile_cap) + kDumpBytes;
}

// What a unit needs to know, worked out once per CTA.
struct unit_info {
	conn_desc const* C;
	int lo, width;            // the tile's targets [lo, lo + width) (local indices)
	std::uint32_t* out;       // counts[slot(step + delay)] + lo
	std::int32_t const* ids0; // the step's slot of the source population's spike ring
	unsigned const* gp;       // run_ptr + tile index: run of source i = groups [gp[i * tiles], gp[i * tiles + 1])
	long long const* tile_ptr; // plain connections: tile_ptr + tile index, stride tiles + 1
	int stride;               // tiles (packed) / tiles + 1 (plain)
	unsigned total;           // spikes of the step (all ranks)
	long long ring_slot;
	// split launches (an item = one round of a unit):
	unsigned b0, b1;          // the item's batches [b0, b1) of the step's spike list
	unsigned round, rounds;   // which of the unit's rounds this is
	unsigned* flag;           // set to the window's epoch once round 0 has stored the unit's counters
};

// One warp's pipeline over its share of a unit: the batches b = first, first + step, ... of the
// step's spike list (a batch = 32 consecutive spikes = 32 runs, padded with empty runs).
// In flight at any time:
//   * registers: the spike ids of the batch after next, the run pointers of the next batch;
//   * shared memory: the descriptors of the current and the next batch;
//   * registers: the entries of the next kRing runs on their way from HBM (one 16-byte load per
//     lane and run), and the run being counted.
// The loop over a half batch is fully unrolled, so the kRing landing slots are plain registers.
// One call counts the batches of one round (at most kRoundBatches, so no u8 counter can wrap).
struct unit_walker {
	tiles_args const& a;
	unit_info const& U;
	unsigned char* smem; // this warp's
	int lane;
	unsigned char* cnt;
	int4 const* stream;  // the connection's packed stream
	unsigned p_g0, p_ng; // this lane's run of the batch whose descriptors are written next
	std::int32_t id_next; // this lane's spike of the batch after that

	__device__ __forceinline__ std::int32_t spike_id(unsigned q) const { return q < U.total ? U.ids0[q] : 0; } // source neuron (0 past the end)
	__device__ __forceinline__ void load_ptrs(unsigned q, std::int32_t id) {
		p_g0 = 0, p_ng = 0;
		if (q < U.total) {
			unsigned const* p = U.gp + static_cast<long long>(id) * U.stride;
			p_g0              = p[0];
			p_ng              = p[1] - p_g0;
		}
	}
	// publish the descriptors of the i-th batch of this warp (from p_g0 / p_ng), then start
	// fetching the pointers of batch `b_next` and the ids of batch `b_next + step`
	__device__ __forceinline__ void write_desc(unsigned i, unsigned b_next, unsigned step) {
		reinterpret_cast<run_desc*>(smem)[(i & 1) * 32 + lane] = run_desc{p_g0, p_ng};
		unsigned const q = b_next * 32 + lane;
		load_ptrs(q, id_next);
		id_next = spike_id(q + step * 32);
		__syncwarp();
	}

	// start fetching the run described by `d`; returns this lane's group of it (in flight)
	__device__ __forceinline__ int4 issue(run_desc const* d) {
		run_desc const r = *d;
		int4 const* g    = stream + r.g0 + lane;
		int4 v           = make_int4(-1, 0, 0, 0);
		if (static_cast<unsigned>(lane) < r.ng)
			v = ldg_stream(g);
		if (r.ng > 32) // run longer than one warp-wide load (rare: tiles are sized for ~100 entries):
			tally_rest(cnt, g, r.ng, lane); // count the rest right away, between two other runs' turns
		return v;
	}

	// batches first, first + step, ... (< nbatch)
	__device__ __forceinline__ void run(unsigned first, unsigned step, unsigned nbatch) {
		cnt    = smem + kDescBytes;
		stream = reinterpret_cast<int4 const*>(U.C->packed);

		zero_tile(cnt, a.tile_cap, (U.width + 127) & ~127, lane);
		unsigned const mine = first < nbatch ? (nbatch - first + step - 1) / step : 0; // batches of this warp
		if (mine > 0) {
			id_next = spike_id(first * 32 + lane);
			load_ptrs(first * 32 + lane, id_next);
			id_next = spike_id((first + step) * 32 + lane);

			run_desc const* const desc = reinterpret_cast<run_desc const*>(smem);
			write_desc(0, first + step, step);
			int4 v[kRing];
#pragma unroll
			for (int j = 0; j < kRing; j++)
				v[j] = issue(desc + j);
			// half batch h: count runs [16h, 16h + 16) of this warp's sequence while fetching [16h + 16, 16h + 32)
			for (unsigned h = 0; h < 2 * mine; h++) {
				if (h & 1) {
					unsigned const i = (h + 1) >> 1; // its descriptors come from batch first + i * step
					write_desc(i, first + (i + 1) * step, step);
				}
				run_desc const* const id = desc + (((h + 1) >> 1) & 1) * 32 + ((h + 1) & 1) * 16;
				bool const more          = h + 1 < 2 * mine;
#pragma unroll
				for (int j = 0; j < kRing; j++) {
					tally(cnt, v[j]);
					__syncwarp();
					// nothing is fetched behind the call's last half batch: the descriptors there belong to the
					// unit's next round (issue() would count the tails of its long runs into this one)
					if (more)
						v[j] = issue(id + j);
				}
			}
		}
	}
};

// Rare path, out of line: connections whose entries are plain columns (rows that may repeat a
// target: adj_list multapses).  One global atomic per event into the tile's (zeroed) range.
__device__ __noinline__ void walk_plain(tiles_args const& a, unit_info const& U, int lane, int warp) {
	long long ev = 0;
	for (unsigned j = warp; j < U.total; j += kWarps) {
		long long const id  = U.ids0[j];
		long long const* tp = U.tile_ptr + id * U.stride;
		long long const beg = tp[0], end = tp[1];
		for (long long e = beg + lane; e < end; e += 32)
			atomicAdd(U.out + (U.C->neighbors[e] - U.lo), 1u);
		ev += end - beg;
	}
	if (lane == 0 && ev)
		atomicAdd(a.stats + 0, static_cast<unsigned long long>(ev));
}

constexpr unsigned kPerRound = kWarps * kRoundBatches; // batches a CTA counts between two merges

__device__ __forceinline__ unsigned ld_acquire(unsigned const* p) {
	unsigned v;
	asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
	asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Split launches: plan[c * nsteps + s] = first item of (connection c, step s), plan[nconns * nsteps] = items.
// (connection, step) holds tiles x rounds items, a unit's rounds next to each other, round 0 first.
__global__ void __launch_bounds__(256) plan_items(tiles_args a) {
	using scan_t = cub::BlockScan<unsigned, 256>;
	__shared__ typename scan_t::TempStorage tmp;
	__shared__ unsigned running;
	int const ncs = a.nconns * a.nsteps;
	if (threadIdx.x == 0)
		running = 0;
	__syncthreads();
	for (int base = 0; base < ncs; base += 256) {
		int const j = base + threadIdx.x;
		unsigned v  = 0;
		if (j < ncs) {
			conn_desc const& C = a.conns[j / a.nsteps];
			long long const t  = a.t0 + j % a.nsteps;
			v                  = static_cast<unsigned>(C.tiles) * rounds_of(C.arranged != 0, C.ring_cnt[(t % a.ring) * C.cnt_stride], a.round_batches);
		}
		unsigned ex, agg;
		scan_t(tmp).ExclusiveSum(v, ex, agg);
		unsigned const r0 = running;
		if (j < ncs)
			a.plan[j] = r0 + ex;
		__syncthreads();
		if (threadIdx.x == 0)
			running = r0 + agg;
		__syncthreads();
	}
	if (threadIdx.x == 0)
		a.plan[ncs] = running;
}

// kSplit = false: a CTA claims whole units and loops over their rounds.
// kSplit = true:  a CTA claims single rounds (plan_items); round 0 of a unit stores the counters and
//                 publishes the unit's flag, later rounds wait for the flag and add with atomics.  The
//                 wait cannot deadlock: round 0 has a lower ticket, so a running CTA holds it, and
//                 round 0 never waits.  For windows with few, long units (a rank of a multi-GPU run).
template <bool kSplit>
__global__ void __launch_bounds__(kWarps * 32, kCtasPerSm) deliver_tiles(tiles_args a) {
	extern __shared__ uint4 smem4[];
	__shared__ unit_info U;
	__shared__ unsigned claimed;
	int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	size_t const wbytes  = warp_smem(a.tile_cap);
	unsigned char* smem  = reinterpret_cast<unsigned char*>(smem4) + warp * wbytes;
	unsigned const units = kSplit ? a.plan[a.nconns * a.nsteps] : static_cast<unsigned>(a.total_tiles) * a.nsteps;
	int cs = 0; // thread 0, split launches: the (connection, step) of the last ticket; tickets only grow
	for (;;) {
		if (threadIdx.x == 0) {
			unsigned const u = atomicAdd(a.work, 1u);
			claimed          = u;
			if (u < units) {
				int c = 0, s, k;
				unsigned r = 0;
				if constexpr (kSplit) {
					while (a.plan[cs + 1] <= u)
						cs++;
					c = cs / a.nsteps;
					s = cs % a.nsteps;
				} else {
					while (c + 1 < a.nconns && static_cast<unsigned>(a.conns[c + 1].tile_prefix) * a.nsteps <= u)
						c++;
				}
				conn_desc const& C = a.conns[c];
				if constexpr (kSplit) {
					unsigned const total = C.ring_cnt[((a.t0 + s) % a.ring) * C.cnt_stride];
					item_pos const it    = locate_item(u - a.plan[cs], C.arranged != 0, total, a.round_batches);
					k                    = static_cast<int>(it.tile);
					r                    = it.round;
					U.round              = it.round;
					U.rounds             = it.rounds;
					U.b0                 = it.b0;
					U.b1                 = it.b1;
					U.flag                = a.unit_flag + (static_cast<unsigned>(C.tile_prefix) * a.nsteps + static_cast<unsigned>(s) * C.tiles + k);
				} else {
					unsigned const local = u - static_cast<unsigned>(C.tile_prefix) * a.nsteps;
					s                    = static_cast<int>(local / C.tiles);
					k                    = static_cast<int>(local % C.tiles);
				}
				long long const t = a.t0 + s;
				U.C         = &C;
				U.lo        = k * C.tile;
				U.width     = static_cast<int>(min(static_cast<long long>(C.tile), C.n_dst - U.lo));
				U.out       = C.counts + ((t + C.delay) % C.cring) * C.cstride + U.lo;
				U.ring_slot = t % a.ring;
				U.ids0      = C.ring_ids + U.ring_slot * C.ring_cap;
				U.gp        = C.run_ptr + k;
				U.tile_ptr  = C.tile_ptr + k;
				U.stride    = C.arranged ? C.tiles : C.tiles + 1;
				unsigned const total = C.ring_cnt[U.ring_slot * C.cnt_stride];
				U.total              = total;
				if (k == 0 && r == 0 && total)
					atomicAdd(a.stats + 1, static_cast<unsigned long long>(total));
			}
		}
		__syncthreads();
		if (claimed >= units)
			break;
		unsigned const nbatch = (U.total + 31) / 32;
		int const words       = (U.width + 3) / 4; // 4 targets per 32-bit word of u8 counters
		if (!U.C->arranged) {
			for (int i = threadIdx.x; i < words * 4; i += kWarps * 32)
				U.out[i] = 0;
			__threadfence_block();
			__syncthreads();
			walk_plain(a, U, lane, warp);
			__syncthreads();
			continue;
		}
		constexpr unsigned per_round = kPerRound;
		unsigned const rounds        = kSplit ? 1u : max(1u, (nbatch + per_round - 1) / per_round);
		for (unsigned round = 0; round < rounds; round++) {
			unsigned const b0 = kSplit ? U.b0 : round * per_round;
			unit_walker w{a, U, smem, lane};
			w.run(b0 + warp, kWarps, kSplit ? U.b1 : min(nbatch, b0 + per_round));
			if constexpr (kSplit)
				if (U.round > 0 && threadIdx.x == 0) // round 0 has stored the unit's counters?
					while (ld_acquire(U.flag) != a.epoch)
						__nanosleep(64);
			__syncthreads();
			// add the warps' arrays and store / accumulate: this CTA is the only writer of the range.
			// Word wd of array A holds targets 4 wd .. 4 wd + 3; their B counters sit in the same
			// 128-byte row r = wd / 32, rotated by r words.  The sum of all counters is the number
			// of Syn::deliver invocations this round stands for.
			uint4* o    = reinterpret_cast<uint4*>(U.out);
			unsigned ev = 0;
			for (int wd = threadIdx.x; wd < words; wd += kWarps * 32) {
				int const r = wd >> 5;
				int const wb = (wd & ~31) | ((wd + r) & 31);
				unsigned even = 0, odd = 0; // two 16-bit lanes each: targets (0, 2) and (1, 3)
#pragma unroll
				for (int w2 = 0; w2 < kWarps; w2++) {
					unsigned char const* base = reinterpret_cast<unsigned char const*>(smem4) + w2 * wbytes + kDescBytes;
					unsigned const ca = reinterpret_cast<unsigned const*>(base)[wd];
					unsigned const cb = reinterpret_cast<unsigned const*>(base + a.tile_cap)[wb];
					even += (ca & 0x00ff00ffu) + (cb & 0x00ff00ffu);
					odd += ((ca >> 8) & 0x00ff00ffu) + ((cb >> 8) & 0x00ff00ffu);
				}
				uint4 c = make_uint4(even & 0xffffu, odd & 0xffffu, even >> 16, odd >> 16);
				ev += c.x + c.y + c.z + c.w;
				if constexpr (kSplit) {
					if (U.round == 0)
						o[wd] = c;
					else {
						if (c.x) atomicAdd(U.out + 4 * wd + 0, c.x);
						if (c.y) atomicAdd(U.out + 4 * wd + 1, c.y);
						if (c.z) atomicAdd(U.out + 4 * wd + 2, c.z);
						if (c.w) atomicAdd(U.out + 4 * wd + 3, c.w);
					}
				} else {
					if (round) {
						uint4 const prev = o[wd];
						c.x += prev.x, c.y += prev.y, c.z += prev.z, c.w += prev.w;
					}
					o[wd] = c;
				}
			}
			for (int off = 16; off; off >>= 1)
				ev += __shfl_xor_sync(kFull, ev, off);
			if (lane == 0 && ev)
				atomicAdd(a.stats + 0, static_cast<unsigned long long>(ev));
			if constexpr (kSplit)
				if (U.round == 0 && U.rounds > 1)
					__threadfence(); // the counters before the flag
			__syncthreads();
			if constexpr (kSplit)
				if (U.round == 0 && U.rounds > 1 && threadIdx.x == 0)
					st_release(U.flag, a.epoch);
		}
	}
}

// ---- pack_runs / unpack_rows ---------------------------------------------------------------------
// Counter addresses of local target t of a tile (t < cap, cap a multiple of 128):
//   array A: t                       (bank (t >> 2) & 31)
//   array B: cap + rot(t),  rot(t) = t with its bank field rotated by the 128-byte row number
//                                      (bank ((t >> 2) + (t >> 7)) & 31)
//   dump:    2 cap + 4 bank + byte   (never read back)
// Two targets that share a bank in A never share one in B (for tiles of <= 32 rows), which is what
// makes the 2-choice balancing effective.
// rot_fwd / rot_inv: deliver_plan.h

// groups of every run: run_ptr[id] = ceil(len / 4) (scanned afterwards); id = row * tiles + k
__global__ void __launch_bounds__(256) run_groups_kernel(long long const* tile_ptr, long long n_runs, int tiles, unsigned* run_ptr) {
	long long const id = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (id > n_runs)
		return;
	unsigned g = 0;
	if (id < n_runs) {
		long long const row = id / tiles;
		int const k         = static_cast<int>(id % tiles);
		g = static_cast<unsigned>((tile_ptr[row * (tiles + 1) + k + 1] - tile_ptr[row * (tiles + 1) + k] + 3) >> 2);
	}
	run_ptr[id] = g;
}

// One thread per run.  A run is packed in chunks of <= 128 entries = 32 groups: the entry at
// position p of a chunk is counted by lane p >> 2 in instruction p & 3, together with the entries
// at p +- 4, +- 8, ... — those must fall into different banks.
__global__ void __launch_bounds__(128) pack_kernel(std::int32_t const* nb, long long const* tile_ptr, unsigned const* run_ptr,
                                                   long long n_runs, int tiles, int tile, int cap, std::int32_t* packed) {
	long long const id = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (id >= n_runs)
		return;
	long long const row = id / tiles;
	int const k         = static_cast<int>(id % tiles);
	long long const beg = tile_ptr[row * (tiles + 1) + k], end = tile_ptr[row * (tiles + 1) + k + 1];
	int const lo        = k * tile;
	std::int32_t* out   = packed + static_cast<long long>(run_ptr[id]) * 4;
	unsigned short t[128];
	unsigned char bank[128]; // chosen bank | array << 7
	unsigned char order[128];
	unsigned char load[32], first[33];
	for (long long g0 = beg; g0 < end; g0 += 128, out += 128) {
		int const m     = static_cast<int>(end - g0 < 128 ? end - g0 : 128);
		int const slots = (m + 3) & ~3;
		for (int b = 0; b < 32; b++)
			load[b] = 0;
		// 2-choice greedy, then two passes that move entries out of banks more than one fuller than their alternative
		for (int j = 0; j < m; j++) {
			int const tt = __ldg(nb + g0 + j) - lo;
			t[j]         = static_cast<unsigned short>(tt);
			int const bA = (tt >> 2) & 31, bB = ((tt >> 2) + (tt >> 7)) & 31;
			bool const pickB = load[bB] < load[bA];
			int const b      = pickB ? bB : bA;
			load[b]++;
			bank[j] = static_cast<unsigned char>(b | (pickB ? 0x80 : 0));
		}
		for (int pass = 0; pass < 2; pass++)
			for (int j = 0; j < m; j++) {
				int const tt  = t[j];
				int const bA  = (tt >> 2) & 31, bB = ((tt >> 2) + (tt >> 7)) & 31;
				bool const inB = (bank[j] & 0x80) != 0;
				int const cur = inB ? bB : bA, alt = inB ? bA : bB;
				if (load[cur] > load[alt] + 1) {
					load[cur]--;
					load[alt]++;
					bank[j] = static_cast<unsigned char>(alt | (inB ? 0 : 0x80));
				}
			}
		// entries grouped by bank
		int maxload = 0, run = 0;
		for (int b = 0; b < 32; b++) {
			first[b] = static_cast<unsigned char>(run);
			run += load[b];
			maxload = load[b] > maxload ? load[b] : maxload;
		}
		first[32] = static_cast<unsigned char>(run);
		for (int b = 0; b < 32; b++)
			load[b] = 0;
		for (int j = 0; j < m; j++) {
			int const b               = bank[j] & 31;
			order[first[b] + load[b]] = static_cast<unsigned char>(j);
			load[b]++;
		}
		// the four instructions of the chunk: free positions, the next free position, the banks in use
		int rem[4], next[4];
		unsigned used_banks[4];
		for (int c = 0; c < 4; c++) {
			next[c]       = c;
			rem[c]        = slots >> 2;
			used_banks[c] = 0;
		}
		// fullest banks first; the entries of one bank go to different instructions, the emptiest first
		for (int L = maxload; L >= 1; L--)
			for (int b = 0; b < 32; b++) {
				if (load[b] != L)
					continue;
				unsigned used = 0;
				for (int i = 0; i < L; i++) {
					int best = -1;
					for (int c = 0; c < 4; c++)
						if (!((used >> c) & 1) && rem[c] > 0 && (best < 0 || rem[c] > rem[best]))
							best = c;
					if (best < 0) { // more entries than instructions left for this bank: conflicts, still correct
						used = 0;
						for (int c = 0; c < 4; c++)
							if (rem[c] > 0 && (best < 0 || rem[c] > rem[best]))
								best = c;
					}
					used |= 1u << best;
					used_banks[best] |= 1u << b;
					rem[best]--;
					int const j  = order[first[b] + i];
					int const tt = t[j];
					out[next[best]] = (bank[j] & 0x80) ? cap + rot_fwd(tt) : tt;
					next[best] += 4;
				}
			}
		// pad to whole groups with dump addresses in banks the instruction does not use
		for (int c = 0; c < 4; c++)
			for (; rem[c] > 0; rem[c]--, next[c] += 4) {
				unsigned const free_banks = ~used_banks[c];
				int const b               = free_banks ? __ffs(free_banks) - 1 : 0;
				used_banks[c] |= 1u << b;
				out[next[c]] = 2 * cap + 4 * b + c;
			}
	}
}

// One thread per row: decode the counter addresses of the row's runs back to local columns and
// emit them ascending (tiles in order, a bitmap per tile).
__global__ void __launch_bounds__(128) unpack_kernel(std::int32_t const* packed, unsigned const* run_ptr, long long const* offsets,
                                                     long long n_rows, int tiles, int tile, int cap, std::int32_t* out) {
	long long const row = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (row >= n_rows)
		return;
	unsigned present[kTileMax / 32];
	int const words = (tile + 31) / 32;
	long long o     = offsets[row];
	for (int k = 0; k < tiles; k++) {
		for (int i = 0; i < words; i++)
			present[i] = 0;
		long long const beg = static_cast<long long>(run_ptr[row * tiles + k]) * 4, end = static_cast<long long>(run_ptr[row * tiles + k + 1]) * 4;
		for (long long e = beg; e < end; e++) {
			int const v = packed[e];
			if (v >= 2 * cap)
				continue;
			int const t = v < cap ? v : rot_inv(v - cap);
			present[t >> 5] |= 1u << (t & 31);
		}
		for (int i = 0; i < words; i++) {
			unsigned m = present[i];
			while (m) {
				int const bit = __ffs(m) - 1;
				m &= m - 1;
				out[o++] = k * tile + i * 32 + bit;
			}
		}
	}
}

__global__ void __launch_bounds__(256) tile_ptr_kernel(long long const* offsets, std::int32_t const* neighbors, long long src_count,
                                                       int tile, int tiles, long long* tile_ptr) {
	long long const idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (idx >= src_count * (tiles + 1))
		return;
	long long const row = idx / (tiles + 1);
	int const k         = static_cast<int>(idx % (tiles + 1));
	long long lo = offsets[row], hi = offsets[row + 1];
	if (k == tiles)
		lo = hi;
	else if (k > 0) {
		long long const want = static_cast<long long>(k) * tile;
		while (lo < hi) { // lower_bound
			long long const mid = (lo + hi) >> 1;
			if (neighbors[mid] < want)
				lo = mid + 1;
			else
				hi = mid;
		}
	}
	tile_ptr[idx] = lo;
}
}

int build_tile_ptr(void* stream, long long const* offsets, std::int32_t const* neighbors, long long src_count, int tile, int tiles,
                   long long* tile_ptr) {
	long long const n = src_count * (tiles + 1);
	if (n > 0)
		tile_ptr_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(offsets, neighbors, src_count,
		                                                                                                        tile, tiles, tile_ptr);
	return static_cast<int>(cudaGetLastError());
}

int count_groups(void* stream, long long const* tile_ptr, long long src_count, int tiles, unsigned* run_ptr, long long* groups_out) {
	auto st           = static_cast<cudaStream_t>(stream);
	long long const n = src_count * tiles;
	run_groups_kernel<<<static_cast<unsigned>((n + 1 + 255) / 256), 256, 0, st>>>(tile_ptr, n, tiles, run_ptr);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess)
		return static_cast<int>(e);
	void* tmp    = nullptr;
	size_t bytes = 0;
	e = cub::DeviceScan::ExclusiveSum(nullptr, bytes, run_ptr, run_ptr, static_cast<long long>(n + 1), st);
	if (e != cudaSuccess)
		return static_cast<int>(e);
	e = cudaMalloc(&tmp, std::max<size_t>(bytes, 16));
	if (e != cudaSuccess)
		return static_cast<int>(e);
	e = cub::DeviceScan::ExclusiveSum(tmp, bytes, run_ptr, run_ptr, static_cast<long long>(n + 1), st);
	unsigned total = 0;
	if (e == cudaSuccess)
		e = cudaMemcpyAsync(&total, run_ptr + n, sizeof(unsigned), cudaMemcpyDeviceToHost, st);
	if (e == cudaSuccess)
		e = cudaStreamSynchronize(st);
	cudaFree(tmp);
	*groups_out = total;
	return static_cast<int>(e);
}

int pack_runs(void* stream, std::int32_t const* neighbors, long long const* tile_ptr, unsigned const* run_ptr, long long src_count,
              int tile, int tiles, int cap, std::int32_t* packed) {
	long long const n = src_count * tiles;
	if (n > 0)
		pack_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(neighbors, tile_ptr, run_ptr, n,
		                                                                                                 tiles, tile, cap, packed);
	return static_cast<int>(cudaGetLastError());
}

int unpack_rows(void* stream, std::int32_t const* packed, unsigned const* run_ptr, long long const* offsets, long long src_count,
                int tile, int tiles, int cap, std::int32_t* out) {
	if (src_count > 0)
		unpack_kernel<<<static_cast<unsigned>((src_count + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
		    packed, run_ptr, offsets, src_count, tiles, tile, cap, out);
	return static_cast<int>(cudaGetLastError());
}

// experiments (read once): SPICE_DELIVER_CTAS_PER_SM / SPICE_DELIVER_GRID shrink the persistent grid,
// SPICE_DELIVER_SPLIT = 0 / 1 forces whole-unit / single-round work items, SPICE_DELIVER_ROUND = batches of 32
// spikes per single-round item (<= kPerRound)
static int env_int(char const* name, int dflt) {
	char const* e = std::getenv(name);
	return e && *e ? std::atoi(e) : dflt;
}
// Split when a window has fewer than 2.5 units per resident CTA.  Measured on one B200 with the per-rank shapes of
// the weak-scaled Brunel benchmark (tools/rank_shape_probe.py, rates oscillating by +-50 %; single rounds vs whole
// units): 1 rank, 4.4 units per CTA: -6 %; 2 ranks, 3.0: -4 %; 4 ranks, 2.3: +7 %; 8 ranks, 1.5: +8 % (+22 % at +-80 %).
constexpr long long kSplitBelowNum = 5, kSplitBelowDen = 2;

int launch_tiles(void* stream, tiles_args const& a, int device, int* launches) {
	static int const ctas_env = env_int("SPICE_DELIVER_CTAS_PER_SM", 0), grid_env = env_int("SPICE_DELIVER_GRID", 0),
	                 split_env = env_int("SPICE_DELIVER_SPLIT", -1), round_env = env_int("SPICE_DELIVER_ROUND", 0);
	static int blocks_per_sm[64] = {};
	static int sms[64]           = {};
	static int smem_set[64]      = {};
	size_t const smem = static_cast<size_t>(kWarps) * warp_smem(a.tile_cap);
	if (device < 0 || device >= 64)
		return static_cast<int>(cudaErrorInvalidDevice);
	if (smem_set[device] < static_cast<int>(smem)) {
		int nb = 1 << 30;
		for (auto kernel : {deliver_tiles<false>, deliver_tiles<true>}) {
			cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
			if (e != cudaSuccess)
				return static_cast<int>(e);
			e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
			if (e != cudaSuccess)
				return static_cast<int>(e);
			int n = 0;
			e     = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kWarps * 32, smem);
			if (e != cudaSuccess)
				return static_cast<int>(e);
			nb = std::min(nb, n);
		}
		cudaDeviceGetAttribute(&sms[device], cudaDevAttrMultiProcessorCount, device);
		blocks_per_sm[device] = std::max(nb, 1);
		if (ctas_env > 0)
			blocks_per_sm[device] = std::clamp(ctas_env, 1, blocks_per_sm[device]);
		smem_set[device] = static_cast<int>(smem);
	}
	long long const units = static_cast<long long>(a.total_tiles) * a.nsteps;
	if (units <= 0)
		return 0;
	long long const full = static_cast<long long>(sms[device]) * blocks_per_sm[device];
	int grid             = static_cast<int>(std::min<long long>(units, full));
	if (grid_env > 0)
		grid = std::clamp(grid_env, 1, grid);
	// few units per CTA (a rank of a multi-GPU run: many sources, few tiles): hand out single rounds
	bool split = a.plan && a.unit_flag && kSplitBelowDen * units < kSplitBelowNum * full;
	if (split_env >= 0)
		split = split_env != 0 && a.plan && a.unit_flag;
	if (launches)
		*launches = split ? 2 : 1;
	if (split) {
		grid = static_cast<int>(grid_env > 0 ? std::min<long long>(grid_env, full) : full); // items >= units; idle CTAs leave at once
		tiles_args b    = a;
		b.round_batches = round_env > 0 ? std::min<unsigned>(round_env, kPerRound) : kPerRound;
		plan_items<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(b);
		deliver_tiles<true><<<grid, kWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(b);
	} else
		deliver_tiles<false><<<grid, kWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(a);
	return static_cast<int>(cudaGetLastError());
}

int preload() {
	cudaFuncAttributes fa{};
	cudaError_t e = cudaFuncGetAttributes(&fa, deliver_tiles<false>);
	if (e == cudaSuccess)
		e = cudaFuncGetAttributes(&fa, deliver_tiles<true>);
	if (e == cudaSuccess)
		e = cudaFuncGetAttributes(&fa, plan_items);
	return static_cast<int>(e);
}
}
