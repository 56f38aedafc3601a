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
//     a row is one contiguous run of ~p * tile column indices;
//   * a unit = (connection, step of the window, tile).  A CTA of 4 warps owns a unit.  Every
//     warp keeps its own copy of the tile's counters as u16 in 10 KB of shared memory and takes
//     every 4th batch of 32 spikes of the step's spike list; it streams each spiking source's
//     run with one 16-byte load per lane (16 runs in flight, in registers) and counts with
//     non-atomic shared-memory read-modify-writes — inside a warp no barrier and no atomic;
//   * at the end the CTA adds its four copies and stores the tile's counters to
//     counts[slot(step + delay)][tile] with plain vector stores: it is the only writer of that
//     range, and the target's update kernel (the only reader) runs in a later window.  Units
//     are handed out by a global counter to a persistent grid, heaviest connections first.
#include <cuda_runtime.h>

#include <algorithm>

#include "deliver.h"

namespace spice::deliver {
namespace {

constexpr int kWarps       = 4;  // warps per CTA: they share a unit, each with its own copy of the tile
constexpr int kCtasPerSm   = 4;  // register budget: 16 warps per SM
constexpr int kRing        = 16; // runs in flight per warp (16 bytes per lane and run, in registers)
constexpr unsigned kFull   = 0xffffffffu;
constexpr unsigned kU16Max = 65535u;

// count the (up to 4) entries of `v` that lie inside the run: entry i has index e0 + i, valid
// when 0 <= e0 + i < len.  The four targets are distinct, so loads may all precede the stores.
template <bool Atomic>
__device__ __forceinline__ void tally(unsigned short* cnt, int4 v, int e0, int len, int lo) {
	bool const p0 = static_cast<unsigned>(e0) < static_cast<unsigned>(len);
	bool const p1 = static_cast<unsigned>(e0 + 1) < static_cast<unsigned>(len);
	bool const p2 = static_cast<unsigned>(e0 + 2) < static_cast<unsigned>(len);
	bool const p3 = static_cast<unsigned>(e0 + 3) < static_cast<unsigned>(len);
	int const t0 = v.x - lo, t1 = v.y - lo, t2 = v.z - lo, t3 = v.w - lo;
	if constexpr (Atomic) {
		unsigned* w = reinterpret_cast<unsigned*>(cnt);
		if (p0) atomicAdd(w + (t0 >> 1), 1u << ((t0 & 1) * 16));
		if (p1) atomicAdd(w + (t1 >> 1), 1u << ((t1 & 1) * 16));
		if (p2) atomicAdd(w + (t2 >> 1), 1u << ((t2 & 1) * 16));
		if (p3) atomicAdd(w + (t3 >> 1), 1u << ((t3 & 1) * 16));
	} else {
		unsigned short c0 = 0, c1 = 0, c2 = 0, c3 = 0;
		if (p0) c0 = cnt[t0];
		if (p1) c1 = cnt[t1];
		if (p2) c2 = cnt[t2];
		if (p3) c3 = cnt[t3];
		if (p0) cnt[t0] = c0 + 1;
		if (p1) cnt[t1] = c1 + 1;
		if (p2) cnt[t2] = c2 + 1;
		if (p3) cnt[t3] = c3 + 1;
	}
}

__device__ __forceinline__ void zero_tile(uint4* cnt4, int words16, int lane) {
	for (int i = lane; i < words16; i += 32)
		cnt4[i] = make_uint4(0, 0, 0, 0);
	__syncwarp();
}

__device__ __forceinline__ int4 ldg_stream(void const* p) {
	int4 v;
	asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
	return v;
}

// One run (a tile's share of one spiking source's row) as the pipeline sees it.
struct run_desc {
	unsigned long long at; // address of the 16-byte group that holds the run's first entry
	int len;               // entries in the run (0: nothing to do)
	int mis;               // entries of that group that precede the run (0..3)
};

// shared memory of one warp: [descriptors: 2 batches x 32 x 16 B][counters: tile_cap x 2 B]
constexpr int kDescBytes = 2 * 32 * static_cast<int>(sizeof(run_desc));
__host__ __device__ constexpr size_t warp_smem(int tile_cap) {
	return static_cast<size_t>(kDescBytes) + static_cast<size_t>(tile_cap) * sizeof(unsigned short);
}

// What a unit needs to know, worked out once per CTA.
struct unit_info {
	conn_desc const* C;
	int lo, width;            // the tile's targets [lo, lo + width) (local indices)
	std::uint32_t* out;       // counts[slot(step + delay)] + lo
	std::int32_t const* ids0; // the step's slot of the source population's spike ring
	long long const* tile_ptr;
	int stride;               // tiles + 1
	unsigned total;           // spikes of the step (all ranks)
	long long ring_slot;
};

// One warp's pipeline over its share of a unit: the batches b = first, first + step, ... of the
// step's spike list (a batch = 32 consecutive spikes = 32 runs, padded with empty runs).
// In flight at any time:
//   * registers: the spike ids of the batch after next, the tile pointers of the next batch;
//   * shared memory: the descriptors of the current and the next batch;
//   * registers: the column indices of the next kRing runs on their way from HBM (one 16-byte
//     load per lane and run), and the run being counted.
// The loop over a half batch is fully unrolled, so the kRing landing slots are plain registers.
// Guarded = true adds the u16 overflow guard (spill the counters to global memory with atomics
// before any of them can wrap) and uses it for the final flush as well.
template <bool Atomic, bool Guarded>
struct unit_walker {
	tiles_args const& a;
	unit_info const& U;
	unsigned char* smem; // this warp's
	int lane;
	unsigned short* cnt;
	unsigned my_n, my_first;
	long long p_beg; // this lane's run of the batch whose descriptors are written next
	int p_len;
	std::int32_t id_next; // this lane's spike of the batch after that
	long long ev;
	unsigned acc; // runs counted since the counters were last zeroed

	__device__ __forceinline__ std::int32_t spike_id(unsigned q) const { // flat index -> source neuron (0 when q >= total)
		int r       = 0;
		unsigned f0 = 0;
		for (int i = 1; i < a.world; i++) {
			unsigned const f = __shfl_sync(kFull, my_first, i);
			unsigned const n = __shfl_sync(kFull, my_n, i);
			if (q >= f && n > 0) {
				r  = i;
				f0 = f;
			}
		}
		return q < U.total ? U.ids0[U.C->seg_lo[r] + (q - f0)] : 0;
	}
	__device__ __forceinline__ void load_ptrs(unsigned q, std::int32_t id) {
		p_beg = 0, p_len = 0;
		if (q < U.total) {
			long long const* p = U.tile_ptr + static_cast<long long>(id) * U.stride;
			p_beg              = p[0];
			p_len              = static_cast<int>(p[1] - p_beg);
		}
	}
	// add this warp's counters to global memory with atomics and clear them (rare path)
	__device__ __forceinline__ void spill_atomic() {
		__syncwarp();
		for (int i = lane; i < U.width; i += 32) {
			unsigned const c = cnt[i];
			if (c)
				atomicAdd(U.out + i, c);
		}
		__syncwarp();
		zero_tile(reinterpret_cast<uint4*>(cnt), (U.width + 7) / 8, lane);
		acc = 0;
	}

	// publish the descriptors of the i-th batch of this warp (from p_beg / p_len), then start
	// fetching the pointers of batch `b_next` and the ids of batch `b_next + step`
	__device__ __forceinline__ void write_desc(unsigned i, unsigned b_next, unsigned step) {
		if constexpr (Guarded) {
			if (acc + 64 + kRing > kU16Max)
				spill_atomic();
			acc += 32;
		}
		ev += p_len;
		int const mis = static_cast<int>(p_beg & 3);
		run_desc d;
		d.at  = reinterpret_cast<unsigned long long>(U.C->neighbors + (p_beg - mis));
		d.len = p_len;
		d.mis = mis;
		reinterpret_cast<run_desc*>(smem)[(i & 1) * 32 + lane] = d;
		unsigned const q = b_next * 32 + lane;
		load_ptrs(q, id_next);
		id_next = spike_id(q + step * 32);
		__syncwarp();
	}

	// start fetching the run described by `d`; returns this lane's 16 bytes of it (in flight)
	__device__ __forceinline__ int4 issue(run_desc const* d) {
		run_desc const r       = *d;
		int const e0           = lane * 4 - r.mis; // index inside the run of this lane's first entry
		unsigned char const* g = reinterpret_cast<unsigned char const*>(r.at) + lane * 16;
		int4 v                 = make_int4(0, 0, 0, 0);
		if (e0 < r.len)
			v = ldg_stream(g);
		if (128 - r.mis < r.len) { // run longer than one warp-wide load (rare: tiles are sized for ~100 entries):
			__syncwarp();          // count the rest right away, between two other runs' turns
			for (int off = 128; off - r.mis < r.len; off += 128) {
				int4 w = make_int4(0, 0, 0, 0);
				if (e0 + off < r.len)
					w = ldg_stream(g + off * 4);
				tally<Atomic>(cnt, w, e0 + off, r.len, U.lo);
			}
			__syncwarp();
		}
		return v;
	}

	// batches first, first + step, ... (< nbatch)
	__device__ __forceinline__ void run(unsigned first, unsigned step, unsigned nbatch) {
		cnt = reinterpret_cast<unsigned short*>(smem + kDescBytes);
		// the step's spike list: one segment per rank; lane r keeps segment r's start in the flat order
		my_n = 0;
		if (lane < a.world)
			my_n = U.C->ring_cnt[U.ring_slot * a.world + lane];
		my_first = my_n;
		for (int off = 1; off < 32; off <<= 1) {
			unsigned const o = __shfl_up_sync(kFull, my_first, off);
			if (lane >= off)
				my_first += o;
		}
		my_first -= my_n; // exclusive prefix

		zero_tile(reinterpret_cast<uint4*>(cnt), (U.width + 7) / 8, lane);
		ev = 0, acc = 0;
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
				run_desc const* const cd = desc + ((h >> 1) & 1) * 32 + (h & 1) * 16;
				run_desc const* const id = desc + (((h + 1) >> 1) & 1) * 32 + ((h + 1) & 1) * 16;
#pragma unroll
				for (int j = 0; j < kRing; j++) {
					int2 const m = *reinterpret_cast<int2 const*>(&cd[j].len);
					tally<Atomic>(cnt, v[j], lane * 4 - m.y, m.x, U.lo);
					__syncwarp();
					v[j] = issue(id + j);
				}
			}
		}
		if constexpr (Guarded)
			spill_atomic();
		for (int off = 16; off; off >>= 1)
			ev += __shfl_xor_sync(kFull, ev, off);
		if (lane == 0 && ev)
			atomicAdd(a.stats + 0, static_cast<unsigned long long>(ev));
	}
};

// the rare variants (rows with repeated targets; more spikes in a step than u16 counters can
// take) stay out of line so that they do not cost the common one registers
__device__ __noinline__ void walk_slow(tiles_args const& a, unit_info const& U, unsigned char* smem, int lane, unsigned first,
                                       unsigned step, unsigned nbatch) {
	unit_walker<true, true> w{a, U, smem, lane};
	w.run(first, step, nbatch);
}

__global__ void __launch_bounds__(kWarps * 32, kCtasPerSm) deliver_tiles(tiles_args a) {
	extern __shared__ uint4 smem4[];
	__shared__ unit_info U;
	__shared__ unsigned claimed;
	int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	size_t const wbytes  = warp_smem(a.tile_cap);
	unsigned char* smem  = reinterpret_cast<unsigned char*>(smem4) + warp * wbytes;
	unsigned const units = static_cast<unsigned>(a.total_tiles) * a.nsteps;
	for (;;) {
		if (threadIdx.x == 0) {
			unsigned const u = atomicAdd(a.work, 1u);
			claimed          = u;
			if (u < units) {
				int c = 0;
				while (c + 1 < a.nconns && static_cast<unsigned>(a.conns[c + 1].tile_prefix) * a.nsteps <= u)
					c++;
				conn_desc const& C   = a.conns[c];
				unsigned const local = u - static_cast<unsigned>(C.tile_prefix) * a.nsteps;
				int const s = static_cast<int>(local / C.tiles), k = static_cast<int>(local % C.tiles);
				long long const t = a.t0 + s;
				U.C         = &C;
				U.lo        = k * C.tile;
				U.width     = static_cast<int>(min(static_cast<long long>(C.tile), C.n_dst - U.lo));
				U.out       = C.counts + ((t + C.delay) % C.cring) * C.cstride + U.lo;
				U.ring_slot = t % a.ring;
				U.ids0      = C.ring_ids + U.ring_slot * C.ring_cap;
				U.tile_ptr  = C.tile_ptr + k;
				U.stride    = C.tiles + 1;
				unsigned total = 0;
				for (int r = 0; r < a.world; r++)
					total += C.ring_cnt[U.ring_slot * a.world + r];
				U.total = total;
				if (k == 0 && total)
					atomicAdd(a.stats + 1, static_cast<unsigned long long>(total));
			}
		}
		__syncthreads();
		if (claimed >= units)
			break;
		unsigned const nbatch = (U.total + 31) / 32;
		int const words16     = (U.width + 7) / 8;
		// every warp can take at most kU16Max runs before a counter could wrap
		bool const big = (nbatch + kWarps - 1) / kWarps * 32 + 64 + kRing > kU16Max;
		if (U.C->atomic || big) {
			// rare: counters go to global memory with atomics; clear the tile's range first
			for (int i = threadIdx.x; i < words16 * 8; i += kWarps * 32)
				U.out[i] = 0;
			__threadfence();
			__syncthreads();
			walk_slow(a, U, smem, lane, warp, kWarps, nbatch);
			__syncthreads();
		} else {
			unit_walker<false, false> w{a, U, smem, lane};
			w.run(warp, kWarps, nbatch);
			__syncthreads();
			// add the warps' copies and store: this CTA is the only writer of the range
			uint4* o = reinterpret_cast<uint4*>(U.out);
			for (int i = threadIdx.x; i < words16; i += kWarps * 32) {
				uint4 lo4 = make_uint4(0, 0, 0, 0), hi4 = lo4;
#pragma unroll
				for (int w2 = 0; w2 < kWarps; w2++) {
					uint4 const c = reinterpret_cast<uint4 const*>(reinterpret_cast<unsigned char*>(smem4) + w2 * wbytes + kDescBytes)[i];
					lo4.x += c.x & 0xffffu, lo4.y += c.x >> 16, lo4.z += c.y & 0xffffu, lo4.w += c.y >> 16;
					hi4.x += c.z & 0xffffu, hi4.y += c.z >> 16, hi4.z += c.w & 0xffffu, hi4.w += c.w >> 16;
				}
				o[2 * i]     = lo4;
				o[2 * i + 1] = hi4;
			}
			__syncthreads();
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

int launch_tiles(void* stream, tiles_args const& a, int device) {
	static int blocks_per_sm[64] = {};
	static int sms[64]           = {};
	static int smem_set[64]      = {};
	size_t const smem = static_cast<size_t>(kWarps) * warp_smem(a.tile_cap);
	if (device < 0 || device >= 64)
		return static_cast<int>(cudaErrorInvalidDevice);
	if (smem_set[device] < static_cast<int>(smem)) {
		cudaError_t e = cudaFuncSetAttribute(deliver_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
		if (e != cudaSuccess)
			return static_cast<int>(e);
		e = cudaFuncSetAttribute(deliver_tiles, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
		if (e != cudaSuccess)
			return static_cast<int>(e);
		int nb = 0;
		e      = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, deliver_tiles, kWarps * 32, smem);
		if (e != cudaSuccess)
			return static_cast<int>(e);
		cudaDeviceGetAttribute(&sms[device], cudaDevAttrMultiProcessorCount, device);
		blocks_per_sm[device] = std::max(nb, 1);
		smem_set[device]      = static_cast<int>(smem);
	}
	long long const units = static_cast<long long>(a.total_tiles) * a.nsteps;
	if (units <= 0)
		return 0;
	int const grid = static_cast<int>(std::min<long long>(units, static_cast<long long>(sms[device]) * blocks_per_sm[device]));
	deliver_tiles<<<grid, kWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(a);
	return static_cast<int>(cudaGetLastError());
}
}
