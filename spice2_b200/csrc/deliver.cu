// Spike delivery, tiled: the reference's hot loop B (spice/include/spice/detail/synapse_population.h:
// 88,99,118-133 under spice/src/snn.cpp:21-25) — for src in spikes, for dst in row(src): deliver —
// for all connections whose events are integer counts (every stateless synapse), one launch per
// window of steps.
//
// Why not one global atomic per event: a B200 SM retires global reductions at ~1.3 cycles per lane,
// i.e. <= 2e11 events/s for the whole GPU, an eighth of what HBM can stream (4 B of CSR per event).
// Shared-memory reductions are another matter: a warp-wide red.shared.add.u32 whose 32 lanes fall
// into 32 different banks retires in ~1 SM cycle (tools/smem_rmw_bench.cu), 32 events per cycle and
// SM, far above what HBM can feed.
//
// Layout of the work:
//   * the targets of a connection are cut into tiles of <= 5120 neurons; a tile's share of a row
//     is one contiguous RUN of ~p * tile entries (sized for ~100: one 16-byte group per lane);
//   * a unit = (connection, step of the window, tile).  A CTA owns a unit at a time and keeps the
//     tile's event counters in shared memory: TWO u32 counters per target (arrays A and B whose
//     bank assignments differ by a per-row rotation, deliver_plan.h).  Once per connection
//     pack_runs() rewrites every run into the kernel's own stream format: each entry becomes the
//     byte offset of one of its target's two counters, chosen (2-choice balancing) and ordered so
//     that the lanes of one counting instruction fall into 32 different banks, the run padded to
//     whole 16-byte groups with offsets of a dump area.  The counting code therefore needs no
//     length, no alignment and no per-entry predicate: a lane holds a whole group or nothing;
//   * the warps of the CTA split the step's spike list in batches of 32 spikes (one run per lane).
//     A warp fetches its runs with cp.async.bulk (1-D TMA), one copy per run issued by the run's
//     lane, eight runs to a stage whose mbarrier collects the bytes; the landed groups are read
//     back with LDS.128 and counted with four red.shared.add per lane.  Nothing sits on the
//     register scoreboards while it is in flight, and the reductions return nothing, so a warp's
//     instruction stream never waits for its own counting;
//   * everything a warp fetches is a few batches ahead of what it counts — spike ids, then run
//     pointers, then the runs — ACROSS unit boundaries: a CTA's next units are known in advance
//     (deliver_plan.h: its first units are fixed by its index, later ones are claimed from a
//     global counter four units ahead), so the pipeline stays full while the CTA merges a unit;
//   * at the end of a unit the CTA adds A + B per target, stores the tile's range of
//     counts[slot(step + delay)] with 16-byte stores and zeroes the counters on the way: it is the
//     only writer of that range, and the target's update kernel (the only reader) runs in a later
//     window, so nobody zeroes anything in global memory.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>

#include <cub/device/device_scan.cuh>

#include "deliver.h"
#include "deliver_plan.h"

namespace spice::deliver {
namespace {

constexpr int kRuns        = 8;   // runs per pipeline stage
constexpr int kSlotBytes   = 512; // landing slot of one run: the first 32 groups (longer runs: see count_stage)
constexpr int kStageBytes  = kRuns * kSlotBytes;
constexpr int kQuarters    = 32 / kRuns; // stages per batch of 32 runs
constexpr int kLookahead   = kStaticUnits - 1; // units beyond the one being counted whose tickets a warp may read
constexpr unsigned kFull   = 0xffffffffu;
// A stream entry is the ADDRESS of a counter in the CTA's shared-memory window (counter_base() + 4 x the word index):
// red.shared takes it as it is, and it is never zero (see issue()).

__device__ __forceinline__ unsigned smem_u32(void const* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
	asm volatile(
	    "{\n"
	    ".reg .pred p;\n"
	    "WAIT_%=:\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
	    "@p bra DONE_%=;\n"
	    "bra WAIT_%=;\n"
	    "DONE_%=:\n"
	    "}\n" ::"r"(bar), "r"(parity)
	    : "memory");
}
// 1-D bulk copy global -> shared, completion (bytes) on an mbarrier
__device__ __forceinline__ void bulk_g2s(unsigned dst, void const* src, unsigned bytes, unsigned bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
	             : "memory");
}
__device__ __forceinline__ int4 lds128(unsigned addr) {
	int4 v;
	asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
	return v;
}
__device__ __forceinline__ int4 ldg_stream(void const* p) {
	int4 v;
	asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
	return v;
}
// count the 4 entries of a group: the shared-memory addresses of 4 u32 counters
__device__ __forceinline__ void tally(int4 v) {
	asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(v.x) : "memory");
	asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(v.y) : "memory");
	asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(v.z) : "memory");
	asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(v.w) : "memory");
}

__device__ __forceinline__ unsigned long long wall_ns() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}

__device__ unsigned g_zero_pair[2]; // an empty run, for lanes without a spike

// what the unit path needs of a connection (a copy in shared memory; conn_desc stays in global memory)
struct conn_hot {
	char const* stream;
	unsigned const* run_ptr;
	std::uint32_t* counts;
	std::int32_t const* ring_ids;
	long long ring_cap, cstride, n_dst;
	int tiles, tile, delay, cring, arranged, pad;
};

// shared state of a CTA (behind the counters: 2 CTAs of 8 warps fit an SM only while this stays below ~9 KB)
template <int kW, int kStages>
struct alignas(16) cta_state {
	unsigned long long bars[kW][kStages];
	int4 units[kTicketRing];             // the CTA's unit sequence, located: units[seq % kTicketRing] = {connection | step << 8 | round << 16, tile,
	                                     // first spike, spikes}; x < 0: no unit
	unsigned cnts[kMaxCounts];           // spikes of (connection, step): world == 1 the total; else the inclusive prefix over ranks
	int prefix[kMaxConns + 1];           // tile_prefix of the connections, + total_tiles
	conn_hot conns[kMaxConns];           // the launch's connections: nothing on the unit path reads them from global memory
};

// A unit as a warp needs it (warp-uniform)
struct unit_view {
	int c, s, k;
	int r;          // round: launches with few, long units hand every unit out in tiles_args::rounds parts of its spike list
	unsigned base;  // first spike of the round in the step's list
	int cs; // (connection, step) index
	unsigned total;
	bool valid;
};

// One batch of this warp: 32 runs, one per lane — kQuarters stages of kRuns consecutive spikes each
struct batch {
	unsigned g0, ng;    // this lane's run: groups [g0, ng) of the connection's stream (ng == g0: no run)
	unsigned seq;       // the unit (CTA sequence number) it belongs to — warp-uniform
	int c;              // its connection — warp-uniform
	bool valid;         // warp-uniform
};

// Rare path, out of line: connections whose entries are plain columns (rows that may repeat a
// target: adj_list multapses).  One global atomic per event into the tile's (zeroed) range.
template <int kW>
__device__ __noinline__ unsigned long long walk_plain(conn_desc const& C, int k, std::int32_t const* ids0, unsigned total, std::uint32_t* out, int lo,
                                                      int lane, int warp) {
	unsigned long long ev = 0;
	long long const* tp0 = C.tile_ptr + k;
	int const stride     = C.tiles + 1;
	for (unsigned j = warp; j < total; j += kW) {
		long long const id  = ids0[j];
		long long const* tp = tp0 + id * stride;
		long long const beg = tp[0], end = tp[1];
		for (long long e = beg + lane; e < end; e += 32)
			atomicAdd(out + (C.neighbors[e] - lo), 1u);
		ev += lane == 0 ? static_cast<unsigned long long>(end - beg) : 0ull;
	}
	return ev;
}

// The spikes of a unit's step are handed to the CTA's warps in QUARTERS of kRuns consecutive spikes, round robin
// (quarter Q goes to warp Q % kW), so the warps' shares of a unit differ by at most one quarter; a warp's batch b is its
// quarters 4 b .. 4 b + 3: lane l holds spike kRuns * (warp + kW * (4 b + l / kRuns)) + l % kRuns of the list.
template <int kW>
__device__ __forceinline__ unsigned warp_batches(unsigned total, int warp) {
	unsigned const nq = (total + kRuns - 1) / kRuns;                                       // quarters of the unit
	unsigned const mine = nq > static_cast<unsigned>(warp) ? (nq - warp + kW - 1) / kW : 0; // quarters of this warp
	return (mine + kQuarters - 1) / kQuarters;
}

// kBulk = true: a run is fetched by its lane with one cp.async.bulk (1-D TMA) against the stage's mbarrier;
// kBulk = false: by the whole warp with one 16-byte cp.async per lane (LDGSTS), stages = commit groups.  Either way
// the run lands in shared memory and nothing waits on a register scoreboard.
template <int kW, int kStages, bool kBulk>
__global__ void __launch_bounds__(kW * 32, kW <= 8 ? 2 : 1) deliver_units(tiles_args a) {
	static_assert(kStages == 2 || kStages == 4, "the stage slots and mbarrier phases below are static for 2 or 4 stages");
	// dynamic shared memory only, the counters first: they start at the same address in every variant of this kernel,
	// the address the stream entries were built for (counter_base(); checked below)
	extern __shared__ uint4 smem4[];
	int const tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	unsigned const rounds = static_cast<unsigned>(a.rounds);
	unsigned const units  = static_cast<unsigned>(a.total_tiles) * a.nsteps * rounds; // work items: (connection, step, tile, round)
	unsigned const cnt    = smem_u32(smem4);                                  // counters: arrays A, B, dump
	int const cnt_words   = 2 * a.tile_cap + 32;
	cta_state<kW, kStages>& sh = *reinterpret_cast<cta_state<kW, kStages>*>(reinterpret_cast<char*>(smem4) + static_cast<size_t>(cnt_words) * 4);
	unsigned const ring   = cnt + static_cast<unsigned>(cnt_words) * 4 + static_cast<unsigned>(sizeof(cta_state<kW, kStages>)) +
	                        static_cast<unsigned>(warp) * (kStages * kStageBytes);
	unsigned const bar0   = smem_u32(&sh.bars[warp][0]);
	if (cnt != a.cnt_base) { // the streams address another shared-memory layout: count nothing
		if (tid == 0)
			atomicOr(a.error, 1 << 16);
		return;
	}

	// ---- CTA prologue -----------------------------------------------------------------------------------
	int const world = a.world;
	if (a.flags) { // several ranks: every peer's spikes of this window have landed in this rank's ring (NVLink peer stores)
		if (tid < world) {
			volatile unsigned long long const* f = a.flags + tid;
			unsigned long long const start       = wall_ns();
			while (*f < a.seq) {
				if (wall_ns() - start > 10000000000ull) { // 10 s of wall time (%globaltimer: independent of the SM clock): a peer died; do not hang the GPU
					atomicOr(a.error, 1);
					break;
				}
				__nanosleep(100);
			}
			__threadfence_system();
		}
		__syncthreads();
	}
	for (int i = tid; i < a.nconns * a.nsteps; i += kW * 32) {
		conn_desc const& C        = a.conns[i / a.nsteps];
		std::uint32_t const* from = C.ring_cnt + ((a.t0 + i % a.nsteps) % a.ring) * C.cnt_stride;
		unsigned run              = 0;
		for (int r = 0; r < world; r++) {
			run += world > 1 ? *reinterpret_cast<volatile std::uint32_t const*>(from + r) : from[r];
			sh.cnts[i * world + r] = run;
		}
	}
	for (int i = tid; i <= a.nconns; i += kW * 32)
		sh.prefix[i] = i < a.nconns ? a.conns[i].tile_prefix : a.total_tiles;
	for (int i = tid; i < a.nconns; i += kW * 32) {
		conn_desc const& C = a.conns[i];
		sh.conns[i] = conn_hot{reinterpret_cast<char const*>(C.packed), C.run_ptr, C.counts, C.ring_ids, C.ring_cap, C.cstride, C.n_dst,
		                       C.tiles, C.tile, static_cast<int>(C.delay), C.cring, C.arranged, 0};
	}
	for (int i = tid; i < (cnt_words + 3) / 4; i += kW * 32)
		smem4[i] = make_uint4(0, 0, 0, 0);
	if (lane < kStages)
		mbar_init(bar0 + lane * 8, 1);
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	// the first dynamically claimed ticket (unit kStaticUnits of this CTA) is on its way while the first unit runs
	unsigned pending = 0;
	if (tid == 0)
		pending = atomicAdd(a.work, 1u);
	__syncthreads();
	// a ticket is located once, by the thread that publishes it: {connection, step, tile, spikes of the step}
	auto locate = [&](unsigned ticket) {
		if (ticket >= units)
			return make_int4(-1, 0, 0, 0);
		unit_pos const p   = locate_unit(ticket / rounds, sh.prefix, a.nconns, a.nsteps);
		unsigned const r   = ticket % rounds;
		unsigned const tot = sh.cnts[(p.c * a.nsteps + p.s) * world + world - 1];
		// round r of a unit: spikes [tot r / rounds, tot (r + 1) / rounds) of the step's list, cut at multiples of a quarter
		unsigned const lo = r == 0 ? 0u : (static_cast<unsigned>(static_cast<unsigned long long>(tot) * r / rounds) & ~(kRuns - 1u));
		unsigned const hi = r + 1 == rounds ? tot : (static_cast<unsigned>(static_cast<unsigned long long>(tot) * (r + 1) / rounds) & ~(kRuns - 1u));
		return make_int4(p.c | (p.s << 8) | (static_cast<int>(r) << 16), p.k, static_cast<int>(lo), static_cast<int>(hi - lo));
	};
	if (tid < kTicketRing)
		sh.units[tid] = tid < kStaticUnits ? locate(static_ticket(blockIdx.x, gridDim.x, tid)) : make_int4(-1, 0, 0, 0);
	__syncthreads();

	auto view = [&](unsigned seq) {
		int4 const r = sh.units[seq % kTicketRing];
		unit_view v{};
		v.valid = r.x >= 0;
		v.c     = r.x & 0xff;
		v.s     = (r.x >> 8) & 0xff;
		v.r     = r.x >> 16;
		v.k     = r.y;
		v.base  = static_cast<unsigned>(r.z);
		v.total = static_cast<unsigned>(r.w);
		v.cs    = v.c * a.nsteps + v.s;
		return v;
	};

	// ---- per-warp pipeline state -------------------------------------------------------------------------
	unsigned done = 0; // units of this CTA merged so far = the unit being counted
	// cursor: the next batch of this warp whose spike ids have not been requested yet
	unsigned cs_seq = 0, cs_b = 0, cs_total = 0, cs_nbatch = 0, cs_base = 0;
	bool cs_known = false, cs_end = false;
	std::int32_t const* cs_ids = nullptr;
	int cs_cs = 0, cs_c = 0, cs_k = 0;
	// spike ids requested (id_*), run pointers requested (nx2), run pointers landed (nxt: its first quarters are issued while
	// cur is counted), runs being fetched and counted (cur)
	bool id_valid = false;
	unsigned id_seq = 0;
	std::int32_t id_src = 0;
	bool id_ok = false; // this lane's share of the batch is a spike (else: behind the end of the list)
	int id_c = 0, id_k = 0;
	batch nx2{0, 0, 0, 0, false}, nxt{0, 0, 0, 0, false}, cur{0, 0, 0, 0, false};
	unsigned q_cur = 0;        // quarters of cur whose copies have been issued (kStages in the steady state)
	unsigned parity = 0;       // kStages == 4: the mbarriers' phase of the batch being counted
	unsigned long long ev = 0; // Syn::deliver invocations this thread has merged

	auto cursor_next = [&](unsigned& seq_out, unsigned& b_out) -> bool {
		for (;;) {
			if (cs_end || cs_seq > done + kLookahead)
				return false;
			if (!cs_known) {
				unit_view const v = view(cs_seq);
				if (!v.valid) {
					cs_end = true;
					return false;
				}
				conn_hot const& C    = sh.conns[v.c];
				long long const slot = (a.t0 + v.s) % a.ring;
				cs_total  = v.total;
				cs_base   = v.base;
				cs_nbatch = C.arranged ? warp_batches<kW>(v.total, warp) : 0; // plain units are walked at their merge
				cs_ids    = C.ring_ids + slot * C.ring_cap;
				cs_cs     = v.cs;
				cs_c      = v.c;
				cs_k      = v.k;
				cs_known  = true;
			}
			if (cs_b < cs_nbatch) {
				seq_out = cs_seq;
				b_out   = cs_b++;
				return true;
			}
			cs_seq++;
			cs_b     = 0;
			cs_known = false;
		}
	};
	// move every prefetch stage forward by one: cur <- nxt <- nx2 <- (run pointers of the ids requested last time) <- (ids
	// of the cursor's next batch).  Nothing here reads what it has just requested: the values are consumed one call later.
	auto advance = [&]() {
		if (!cur.valid && nxt.valid) {
			cur       = nxt;
			nxt.valid = false;
		}
		if (!nxt.valid && nx2.valid) {
			nxt       = nx2;
			nx2.valid = false;
		}
		if (!nx2.valid && id_valid) {
			nx2.valid = true;
			nx2.seq   = id_seq;
			nx2.c     = id_c;
			// always a load, never a merge with a constant: the two words land in nx2's own registers and nobody waits
			// for them before the next call (a lane without a spike reads a pair of zeros: an empty run)
			unsigned const* p = id_ok ? sh.conns[id_c].run_ptr + (static_cast<long long>(id_src) * sh.conns[id_c].tiles + id_k) : g_zero_pair;
			nx2.g0            = p[0];
			nx2.ng            = p[1]; // the END of the run: issue() subtracts
			id_valid          = false;
		}
		if (!id_valid) {
			unsigned seq, b;
			if (cursor_next(seq, b)) {
				unsigned const q  = kRuns * (warp + kW * (kQuarters * b + lane / kRuns)) + lane % kRuns;
				unsigned const qc = min(q, cs_total - 1); // (a cursor batch exists only in units with spikes)
				id_valid = true;
				id_seq   = seq;
				id_ok    = q < cs_total;
				unsigned const g  = cs_base + qc; // its place in the step's list
				if (world == 1)
					id_src = cs_ids[g];
				else { // spike g of the step: the (g - spikes of the ranks before r)-th of rank r's segment
					unsigned const* pre = sh.cnts + cs_cs * world;
					int r               = 0;
					while (pre[r] <= g)
						r++;
					id_src = *reinterpret_cast<volatile std::int32_t const*>(cs_ids + a.conns[cs_c].seg_lo[r] + (g - (r ? pre[r - 1] : 0u)));
				}
				id_c = cs_c;
				id_k = cs_k;
			}
		}
	};
	// Once per batch.  In the steady state one step fills cur from values requested a whole batch ago; only an empty
	// pipeline (the CTA's first unit, or a cursor that had run into the look-ahead limit) takes the dependent steps, behind
	// a real branch: a move of a value still in flight would wait for it whether it is needed or not.
	auto refill = [&]() {
#pragma unroll 1
		for (int pass = 0; pass < 4; pass++) { // ONE copy of the code: the state keeps its registers, nothing is moved (and waited for) at a join
			advance();
			if (cur.valid || !(nxt.valid || nx2.valid || id_valid))
				break;
		}
	};
	// issue the copies of quarter q of batch b into stage slot `st`
	auto issue = [&](batch const& b, unsigned q, unsigned st) {
		char const* stream = sh.conns[b.c].stream;
		if constexpr (kBulk) {
			unsigned const bar   = bar0 + st * 8;
			bool const mine      = (static_cast<unsigned>(lane) / kRuns) == q;
			unsigned const bytes = mine ? min(b.ng - b.g0, 32u) * 16 : 0;
			unsigned const tot   = __reduce_add_sync(kFull, bytes);
			if (lane == 0)
				mbar_arrive_expect_tx(bar, tot);
			__syncwarp();
			if (bytes)
				bulk_g2s(ring + st * kStageBytes + (lane % kRuns) * kSlotBytes, stream + static_cast<unsigned long long>(b.g0) * 16, bytes, bar);
		} else {
			// every lane copies its group of every run, or zero-fills its place in the run's slot (src-size 0: nothing is
			// read): a stream entry is never zero (an address behind the window's reserved start), so the count tells the two apart without the run's length
			unsigned const dst = ring + st * kStageBytes + lane * 16;
#pragma unroll
			for (int j = 0; j < kRuns; j++) {
				unsigned const g0 = __shfl_sync(kFull, b.g0, q * kRuns + j);
				unsigned const g1 = __shfl_sync(kFull, b.ng, q * kRuns + j);
				unsigned const sz = static_cast<unsigned>(lane) < g1 - g0 ? 16u : 0u;
				asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + j * kSlotBytes),
				             "l"(stream + static_cast<unsigned long long>(g0 + lane) * 16), "r"(sz)
				             : "memory");
			}
			asm volatile("cp.async.commit_group;" ::: "memory");
		}
	};
	// count quarter q of cur, landed in stage slot `st`
	auto count_stage = [&](unsigned q, unsigned st, unsigned ph, unsigned longer) {
		unsigned const base = ring + st * kStageBytes + lane * 16;
		if constexpr (kBulk) {
			mbar_wait(bar0 + st * 8, ph);
			unsigned const mine = cur.ng - cur.g0;
#pragma unroll
			for (int h = 0; h < kRuns; h += 4) { // four runs at a time: their groups are read back, then counted
				int4 v[4];
				unsigned n[4];
#pragma unroll
				for (int j = 0; j < 4; j++) {
					n[j] = __shfl_sync(kFull, mine, q * kRuns + h + j);
					v[j] = lds128(base + (h + j) * kSlotBytes); // (a stale slot where this lane has no group: not counted)
				}
#pragma unroll
				for (int j = 0; j < 4; j++)
					if (static_cast<unsigned>(lane) < n[j])
						tally(v[j]);
			}
		} else {
			// groups complete in order: all but the kStages - 1 youngest have landed (each lane reads back only what it copied)
			asm volatile("cp.async.wait_group %0;" ::"n"(kStages - 1) : "memory");
#pragma unroll
			for (int h = 0; h < kRuns; h += 4) {
				int4 v[4];
#pragma unroll
				for (int j = 0; j < 4; j++)
					v[j] = lds128(base + (h + j) * kSlotBytes);
#pragma unroll
				for (int j = 0; j < 4; j++)
					if (v[j].x)
						tally(v[j]);
			}
		}
		if ((longer >> (q * kRuns)) & ((1u << kRuns) - 1)) { // a run of more than 32 groups in this quarter (rare: tiles are sized for ~25): the rest straight from global memory
			char const* stream  = sh.conns[cur.c].stream;
			unsigned const mine = cur.ng - cur.g0;
#pragma unroll 1
			for (int j = 0; j < kRuns; j++) {
				unsigned const g0 = __shfl_sync(kFull, cur.g0, q * kRuns + j);
				unsigned const nj = __shfl_sync(kFull, mine, q * kRuns + j);
				for (unsigned off = 32 + lane; off < nj; off += 32)
					tally(ldg_stream(stream + static_cast<unsigned long long>(g0 + off) * 16));
			}
		}
		__syncwarp(); // every lane has read the stage's slots: they may be overwritten
	};
	// end of unit `done`: all warps arrive; merge + zero the counters, store the tile's range; publish the next ticket
	auto boundary = [&](unit_view const& U) {
		__syncthreads();
		conn_hot const& C  = sh.conns[U.c];
		int const lo       = U.k * C.tile;
		int const width    = static_cast<int>(min(static_cast<long long>(C.tile), C.n_dst - lo));
		long long const t  = a.t0 + U.s;
		std::uint32_t* out = C.counts + ((t + C.delay) % C.cring) * C.cstride + lo;
		int const quads    = (width + 3) / 4;
		uint4* o           = reinterpret_cast<uint4*>(out);
		if (tid == 0 && U.k == 0 && U.total) // (every round adds its own share of the step's spikes)
			atomicAdd(a.stats + 1, static_cast<unsigned long long>(U.total));
		if (!C.arranged) {
			// plain columns: the whole unit in its round 0 (the rare path is not split)
			if (!a.prezeroed) {
				for (int i = tid; i < quads; i += kW * 32)
					o[i] = make_uint4(0, 0, 0, 0);
				__threadfence_block();
				__syncthreads();
			}
			unsigned const tot = sh.cnts[U.cs * world + world - 1];
			if (tot && U.r == 0) {
				std::int32_t const* ids = C.ring_ids + ((a.t0 + U.s) % a.ring) * C.ring_cap;
				unsigned before         = 0;
				for (int r = 0; r < world; r++) { // one segment per rank (one rank: the whole list)
					unsigned const upto = sh.cnts[U.cs * world + r];
					ev += walk_plain<kW>(a.conns[U.c], U.k, ids + (world > 1 ? a.conns[U.c].seg_lo[r] : 0), upto - before, out, lo, lane, warp);
					before = upto;
				}
			}
		} else if (U.total == 0) {
			if (!a.prezeroed) // (pre-zeroed counters: the consumer clears what it has read, nothing to store)
				for (int i = tid; i < quads; i += kW * 32)
					o[i] = make_uint4(0, 0, 0, 0);
		} else {
			unsigned const cap4 = static_cast<unsigned>(a.tile_cap) * 4;
			for (int i = tid; i < quads; i += kW * 32) {
				int const t0     = 4 * i;
				unsigned const A = cnt + 16u * i;
				int4 ca          = lds128(A);
				asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(A), "r"(0u) : "memory");
				int const sh_    = rotw_amount(t0 >> 5);
				unsigned const B = cnt + cap4 + 4u * (t0 & ~31);
				unsigned cb[4];
#pragma unroll
				for (int j = 0; j < 4; j++) {
					unsigned const at = B + 4u * ((t0 + j + sh_) & 31);
					asm volatile("ld.shared.u32 %0, [%1];" : "=r"(cb[j]) : "r"(at));
					asm volatile("st.shared.u32 [%0], %1;" ::"r"(at), "r"(0u) : "memory");
				}
				uint4 const c = make_uint4(ca.x + cb[0], ca.y + cb[1], ca.z + cb[2], ca.w + cb[3]);
				ev += c.x + c.y + c.z + c.w;
				if (rounds == 1)
					o[i] = c; // the only writer of this range in this window
				else { // one of several rounds: added to counters the consumer left at zero (integer sums: any order, same bits)
					std::uint32_t* at = out + t0;
					if (c.x)
						asm volatile("red.global.add.u32 [%0], %1;" ::"l"(at), "r"(c.x) : "memory");
					if (c.y)
						asm volatile("red.global.add.u32 [%0], %1;" ::"l"(at + 1), "r"(c.y) : "memory");
					if (c.z)
						asm volatile("red.global.add.u32 [%0], %1;" ::"l"(at + 2), "r"(c.z) : "memory");
					if (c.w)
						asm volatile("red.global.add.u32 [%0], %1;" ::"l"(at + 3), "r"(c.w) : "memory");
				}
			}
		}
		if (tid == 0) { // the ticket claimed while this unit ran becomes unit done + kStaticUnits; claim the one after it
			sh.units[(done + kStaticUnits) % kTicketRing] = locate(dynamic_ticket(gridDim.x, pending));
			pending                                         = atomicAdd(a.work, 1u);
		}
		__syncthreads();
		done++;
	};

	// ---- main loop: one iteration per batch of this warp ---------------------------------------------------
	// Steady state: on entry the first kStages quarters of cur are in flight (issued while the previous batch was
	// counted); quarter q is counted from stage slot q % kStages, and the slot is refilled at once with the quarter
	// kStages further on — of cur, or of nxt.  A batch uses every slot 4 / kStages times, so slots and (for two
	// stages) mbarrier phases are compile-time constants.
	for (;;) {
		refill();
		if (!cur.valid || cur.seq != done) {
			// nothing left for this warp in unit `done` (its next batch, if any, belongs to a later unit)
			unit_view const U = view(done);
			if (!U.valid)
				break;
			boundary(U);
			continue;
		}
#pragma unroll
		for (int q = 0; q < kStages; q++) // after a pipeline bubble only
			if (q_cur <= static_cast<unsigned>(q))
				issue(cur, q, q);
		unsigned const longer = __ballot_sync(kFull, cur.ng - cur.g0 > 32u); // lanes whose run has more than 32 groups
#pragma unroll
		for (int q = 0; q < kQuarters; q++) {
			count_stage(q, q % kStages, kStages == 2 ? (q >> 1) & 1 : parity, longer);
			if (q + kStages < kQuarters)
				issue(cur, q + kStages, q % kStages);
			else if (nxt.valid)
				issue(nxt, q + kStages - kQuarters, q % kStages);
			else if constexpr (!kBulk) // keep the stage being counted kStages - 1 commit groups behind the youngest
				asm volatile("cp.async.commit_group;" ::: "memory");
		}
		parity ^= 1;
		q_cur     = nxt.valid ? kStages : 0;
		cur.valid = false;
	}
	for (int off = 16; off; off >>= 1)
		ev += __shfl_xor_sync(kFull, ev, off);
	if (lane == 0 && ev)
		atomicAdd(a.stats + 0, ev);
}

// ---- pack_runs / unpack_rows ---------------------------------------------------------------------
// Counter words of local target t of a tile (t < cap, cap a multiple of 128), deliver_plan.h:
//   array A: word t                  (bank t & 31)
//   array B: word cap + rotw_fwd(t)  (bank (t + rotw_amount(t / 32)) & 31)
//   dump:    word 2 cap + bank       (never read back)
// A stream entry is 4 x the word index (the byte offset red.shared takes).

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
                                                   long long n_runs, int tiles, int tile, int cap, int base, std::int32_t* packed) {
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
			int const bA = tt & 31, bB = (tt + rotw_amount(tt >> 5)) & 31;
			bool const pickB = load[bB] < load[bA];
			int const b      = pickB ? bB : bA;
			load[b]++;
			bank[j] = static_cast<unsigned char>(b | (pickB ? 0x80 : 0));
		}
		for (int pass = 0; pass < 2; pass++)
			for (int j = 0; j < m; j++) {
				int const tt  = t[j];
				int const bA  = tt & 31, bB = (tt + rotw_amount(tt >> 5)) & 31;
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
					out[next[best]] = base + 4 * ((bank[j] & 0x80) ? cap + rotw_fwd(tt) : tt);
					next[best] += 4;
				}
			}
		// pad to whole groups with dump addresses in banks the instruction does not use
		for (int c = 0; c < 4; c++)
			for (; rem[c] > 0; rem[c]--, next[c] += 4) {
				unsigned const free_banks = ~used_banks[c];
				int const b               = free_banks ? __ffs(free_banks) - 1 : 0;
				used_banks[c] |= 1u << b;
				out[next[c]] = base + 4 * (2 * cap + b);
			}
	}
}

// One thread per row: decode the counter addresses of the row's runs back to local columns and
// emit them ascending (tiles in order, a bitmap per tile).
__global__ void __launch_bounds__(128) unpack_kernel(std::int32_t const* packed, unsigned const* run_ptr, long long const* offsets,
                                                     long long n_rows, int tiles, int tile, int cap, int base, std::int32_t* out) {
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
			int const v = (packed[e] - base) >> 2;
			if (v >= 2 * cap)
				continue;
			int const t = v < cap ? v : rotw_inv(v - cap);
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

namespace {
// where a kernel without static shared memory finds its dynamic shared memory (the window's reserved start lies before it)
__global__ void smem_base_kernel(unsigned* out) {
	extern __shared__ uint4 probe4[];
	*out = smem_u32(probe4);
}
}

int counter_base(void* stream, int* base_out) {
	static std::mutex m;
	static int cached[64] = {};
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess || dev < 0 || dev >= 64)
		return static_cast<int>(e != cudaSuccess ? e : cudaErrorInvalidDevice);
	std::lock_guard<std::mutex> lock(m);
	if (!cached[dev]) {
		auto st       = static_cast<cudaStream_t>(stream);
		unsigned* d   = nullptr;
		unsigned base = 0;
		e = cudaMalloc(&d, sizeof(unsigned));
		if (e != cudaSuccess)
			return static_cast<int>(e);
		smem_base_kernel<<<1, 1, 16, st>>>(d);
		e = cudaGetLastError();
		if (e == cudaSuccess)
			e = cudaMemcpyAsync(&base, d, sizeof(unsigned), cudaMemcpyDeviceToHost, st);
		if (e == cudaSuccess)
			e = cudaStreamSynchronize(st);
		cudaFree(d);
		if (e != cudaSuccess)
			return static_cast<int>(e);
		if (base == 0 || base > 0x10000)
			return static_cast<int>(cudaErrorUnknown);
		cached[dev] = static_cast<int>(base);
	}
	*base_out = cached[dev];
	return 0;
}

int pack_runs(void* stream, std::int32_t const* neighbors, long long const* tile_ptr, unsigned const* run_ptr, long long src_count,
              int tile, int tiles, int cap, std::int32_t* packed) {
	long long const n = src_count * tiles;
	int base          = 0;
	if (int const e = counter_base(stream, &base))
		return e;
	if (n > 0)
		pack_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(neighbors, tile_ptr, run_ptr, n,
		                                                                                                 tiles, tile, cap, base, packed);
	return static_cast<int>(cudaGetLastError());
}

int unpack_rows(void* stream, std::int32_t const* packed, unsigned const* run_ptr, long long const* offsets, long long src_count,
                int tile, int tiles, int cap, std::int32_t* out) {
	int base = 0;
	if (int const e = counter_base(stream, &base))
		return e;
	if (src_count > 0)
		unpack_kernel<<<static_cast<unsigned>((src_count + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
		    packed, run_ptr, offsets, src_count, tiles, tile, cap, base, out);
	return static_cast<int>(cudaGetLastError());
}

// experiments (read once): SPICE_DELIVER_WARPS = 8 / 16 picks the CTA shape, SPICE_DELIVER_GRID shrinks the persistent grid
static int env_int(char const* name, int dflt) {
	char const* e = std::getenv(name);
	return e && *e ? std::atoi(e) : dflt;
}

namespace {
size_t cta_smem(int tile_cap, int warps, int stages) {
	size_t const state = warps == 8 ? sizeof(cta_state<8, 2>) : sizeof(cta_state<16, 2>);
	return static_cast<size_t>(2 * tile_cap + 32) * 4 + state + static_cast<size_t>(warps) * stages * kStageBytes;
}
using kernel_t = void (*)(tiles_args);
struct shape {
	kernel_t kernel;
	int warps, stages;
};
// two CTAs of 8 warps per SM (one merges while the other streams), or one of 16 warps: windows with few, long units
// (a rank of a multi-GPU run sees every source but few target tiles)
shape const kShapes[4] = {{deliver_units<8, 2, true>, 8, 2}, {deliver_units<16, 2, true>, 16, 2},
                          {deliver_units<8, 2, false>, 8, 2}, {deliver_units<16, 2, false>, 16, 2}};
}

int launch_tiles(void* stream, tiles_args const& a, int device, int* launches) {
	static int const warps_env = env_int("SPICE_DELIVER_WARPS", 0), grid_env = env_int("SPICE_DELIVER_GRID", 0), rounds_env = env_int("SPICE_DELIVER_ROUNDS", 0),
	                 path_env = env_int("SPICE_DELIVER_PATH", 0); // 0: cp.async.bulk + mbarrier (measured faster: 0.658 vs 0.625 of the roofline), 1: cp.async with zero fill
	static int blocks_per_sm[64][4] = {};
	static int sms[64]              = {};
	static int smem_set[64]         = {};
	if (device < 0 || device >= 64)
		return static_cast<int>(cudaErrorInvalidDevice);
	if (a.nconns > kMaxConns || a.nsteps > spice::detail::kMaxWindow || a.nconns * a.nsteps * a.world > kMaxCounts)
		return static_cast<int>(cudaErrorInvalidValue);
	if (smem_set[device] < a.tile_cap) {
		for (int i = 0; i < 4; i++) {
			size_t const smem = cta_smem(a.tile_cap, kShapes[i].warps, kShapes[i].stages);
			cudaError_t e = cudaFuncSetAttribute(kShapes[i].kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
			if (e != cudaSuccess)
				return static_cast<int>(e);
			e = cudaFuncSetAttribute(kShapes[i].kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
			if (e != cudaSuccess)
				return static_cast<int>(e);
			int n = 0;
			e     = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kShapes[i].kernel, kShapes[i].warps * 32, smem);
			if (e != cudaSuccess)
				return static_cast<int>(e);
			blocks_per_sm[device][i] = std::max(n, 1);
		}
		cudaDeviceGetAttribute(&sms[device], cudaDevAttrMultiProcessorCount, device);
		smem_set[device] = a.tile_cap;
	}
	long long units = static_cast<long long>(a.total_tiles) * a.nsteps;
	if (units <= 0)
		return 0;
	// Few, long units (a rank of a multi-GPU run sees every source but only its share of the targets): measured on the
	// 8-rank shape (tools/rank_shape_probe.py 8 300 0.5, profiles/probe_r02_rank_shapes.txt) two CTAs of 8 warps on whole
	// units deliver in 190 us per window, one CTA of 16 warps in 207 us, units handed out in 2 / 3 / 4 rounds that add to
	// pre-zeroed counters in 211 / 209 / 215 us: whole units on the 8-warp shape it is.  The other two stay reachable for
	// experiments (SPICE_DELIVER_WARPS=16; SPICE_PREZEROED=1 SPICE_DELIVER_ROUNDS=n).
	int rounds = 1;
	int which  = 0;
	if (rounds_env > 0 && a.prezeroed)
		rounds = std::clamp(rounds_env, 1, 8);
	units *= rounds;
	if (warps_env == 8 || warps_env == 16)
		which = warps_env == 16;
	which += path_env == 1 ? 2 : 0;
	shape const& S       = kShapes[which];
	long long const full = static_cast<long long>(sms[device]) * blocks_per_sm[device][which];
	int grid             = static_cast<int>(std::min<long long>(units, full));
	if (grid_env > 0)
		grid = std::clamp(grid_env, 1, grid);
	if (launches)
		*launches = 1;
	tiles_args b = a;
	int base     = 0;
	if (int const e = counter_base(stream, &base))
		return e;
	b.cnt_base = static_cast<unsigned>(base);
	b.rounds   = rounds;
	S.kernel<<<grid, S.warps * 32, cta_smem(a.tile_cap, S.warps, S.stages), static_cast<cudaStream_t>(stream)>>>(b);
	return static_cast<int>(cudaGetLastError());
}

int preload() {
	cudaFuncAttributes fa{};
	cudaError_t e = cudaSuccess;
	for (int i = 0; i < 4 && e == cudaSuccess; i++)
		e = cudaFuncGetAttributes(&fa, kShapes[i].kernel);
	return static_cast<int>(e);
}
}
