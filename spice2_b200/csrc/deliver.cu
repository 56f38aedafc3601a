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
//   * a unit = (connection, step of the window, tile).  A warp owns a unit: it keeps the
//     tile's counters as u16 in its private 10 KB of shared memory, walks the step's spike
//     list, streams each spiking source's run with 16-byte loads (8 runs in flight per lane)
//     and counts with non-atomic shared-memory read-modify-writes — no barrier, no atomic;
//   * at the end the warp stores the tile's counters to counts[slot(step + delay)][tile] with
//     plain vector stores: it is the only writer of that range, and the target's update kernel
//     (the only reader) runs in a later window.  Units are handed out by a global counter to a
//     persistent grid (22 warps per SM), heaviest connections first.
#include <cuda_runtime.h>

#include <algorithm>

#include "deliver.h"

namespace spice::deliver {
namespace {

constexpr int kWarps = 2;  // warps per CTA (each with its own tile)
constexpr int kDepth = 8;  // runs whose first 16-byte load is in flight per lane
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ int4 ldg_stream(int4 const* p) {
	int4 v;
	asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
	return v;
}

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

// counts[out + i] (=|+=) cnt[i] for i < width (rounded up to 8: the row stride is padded)
__device__ __forceinline__ void flush_tile(uint4 const* cnt4, std::uint32_t* out, int width, bool add, int lane) {
	__syncwarp();
	uint4* o = reinterpret_cast<uint4*>(out);
	for (int i = lane; i * 8 < width; i += 32) {
		uint4 const w = cnt4[i];
		uint4 a = make_uint4(w.x & 0xffffu, w.x >> 16, w.y & 0xffffu, w.y >> 16);
		uint4 b = make_uint4(w.z & 0xffffu, w.z >> 16, w.w & 0xffffu, w.w >> 16);
		if (add) {
			uint4 const pa = o[2 * i], pb = o[2 * i + 1];
			a.x += pa.x, a.y += pa.y, a.z += pa.z, a.w += pa.w;
			b.x += pb.x, b.y += pb.y, b.z += pb.z, b.w += pb.w;
		}
		o[2 * i]     = a;
		o[2 * i + 1] = b;
	}
}

template <bool Atomic>
__device__ __forceinline__ void run_unit(tiles_args const& a, conn_desc const& C, int s, int k, unsigned short* cnt, int lane) {
	uint4* cnt4          = reinterpret_cast<uint4*>(cnt);
	int const B          = C.tile;
	int const lo         = k * B;
	int const width      = static_cast<int>(min(static_cast<long long>(B), C.n_dst - lo));
	int const words16    = (width + 7) / 8;
	long long const t    = a.t0 + s;
	long long const slot = t % a.ring;
	std::uint32_t* out   = C.counts + ((t + C.delay) % C.cring) * C.cstride + lo;
	std::int32_t const* nb    = C.neighbors;
	long long const* tile_ptr = C.tile_ptr + k;
	int const stride          = C.tiles + 1;

	zero_tile(cnt4, words16, lane);
	unsigned acc   = 0; // spikes counted since the last flush (u16 counters: flush before 65536)
	bool first     = true;
	long long ev   = 0;
	unsigned total = 0;
	for (int r = 0; r < a.world; r++) {
		unsigned const n        = C.ring_cnt[slot * a.world + r];
		std::int32_t const* ids = C.ring_ids + slot * C.ring_cap + C.seg_lo[r];
		total += n;
		for (unsigned base = 0; base < n; base += 32) {
			if (acc + 32 > 65535u) {
				flush_tile(cnt4, out, width, !first, lane);
				first = false;
				zero_tile(cnt4, words16, lane);
				acc = 0;
			}
			unsigned const j = base + lane;
			long long beg    = 0;
			int len          = 0;
			if (j < n) {
				long long const* p = tile_ptr + static_cast<long long>(ids[j]) * stride;
				beg                = p[0];
				len                = static_cast<int>(p[1] - beg);
			}
			ev += len;
			int const m = static_cast<int>(min(32u, n - base));
			acc += m;
			for (int g = 0; g < m; g += kDepth) {
				int4 v[kDepth];
				int ln[kDepth];
				unsigned mispack = 0; // 2 bits per run: how far its start is from a 16-byte boundary
#pragma unroll
				for (int d = 0; d < kDepth; d++) {
					int const sl      = (g + d) & 31;
					long long const b = __shfl_sync(kFull, beg, sl);
					int l             = __shfl_sync(kFull, len, sl);
					if (g + d >= m)
						l = 0;
					int const mis = static_cast<int>(b & 3);
					ln[d]         = l;
					mispack |= static_cast<unsigned>(mis) << (2 * d);
					v[d] = make_int4(0, 0, 0, 0);
					if (l > 0 && lane * 4 - mis < l) // lane's first entry has index lane*4 - mis inside the run
						v[d] = ldg_stream(reinterpret_cast<int4 const*>(nb + (b - mis)) + lane);
				}
#pragma unroll
				for (int d = 0; d < kDepth; d++) {
					int const l = ln[d];
					if (l == 0)
						continue;
					int const mis = static_cast<int>((mispack >> (2 * d)) & 3u);
					int e0        = lane * 4 - mis;
					tally<Atomic>(cnt, v[d], e0, l, lo);
					if (128 - mis < l) { // run longer than one warp-wide load (rare: tiles are sized for ~100 entries)
						long long const b = __shfl_sync(kFull, beg, (g + d) & 31) - mis;
						for (int off = 128; off - mis < l; off += 128) {
							e0 += 128;
							int4 w = make_int4(0, 0, 0, 0);
							if (e0 < l)
								w = ldg_stream(reinterpret_cast<int4 const*>(nb + b + off) + lane);
							tally<Atomic>(cnt, w, e0, l, lo);
						}
					}
					__syncwarp();
				}
			}
		}
	}
	flush_tile(cnt4, out, width, !first, lane);
	__syncwarp();
	for (int off = 16; off; off >>= 1)
		ev += __shfl_xor_sync(kFull, ev, off);
	if (lane == 0) {
		if (ev)
			atomicAdd(a.stats + 0, static_cast<unsigned long long>(ev));
		if (k == 0 && total)
			atomicAdd(a.stats + 1, static_cast<unsigned long long>(total));
	}
}

// the rare variant (rows with repeated targets) stays out of line so that it does not cost the
// common one registers
__device__ __noinline__ void run_unit_atomic(tiles_args const& a, conn_desc const& C, int s, int k, unsigned short* cnt, int lane) {
	run_unit<true>(a, C, s, k, cnt, lane);
}

__global__ void __launch_bounds__(kWarps * 32, 20 / kWarps) deliver_tiles(tiles_args a) {
	extern __shared__ uint4 smem4[];
	int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned short* cnt  = reinterpret_cast<unsigned short*>(smem4) + static_cast<size_t>(warp) * a.tile_cap;
	unsigned const units = static_cast<unsigned>(a.total_tiles) * a.nsteps;
	for (;;) {
		unsigned u = 0;
		if (lane == 0)
			u = atomicAdd(a.work, 1u);
		u = __shfl_sync(kFull, u, 0);
		if (u >= units)
			break;
		int c = 0;
		while (c + 1 < a.nconns && static_cast<unsigned>(a.conns[c + 1].tile_prefix) * a.nsteps <= u)
			c++;
		conn_desc const& C  = a.conns[c];
		unsigned const local = u - static_cast<unsigned>(C.tile_prefix) * a.nsteps;
		int const s = static_cast<int>(local / C.tiles), k = static_cast<int>(local % C.tiles);
		if (C.atomic)
			run_unit_atomic(a, C, s, k, cnt, lane);
		else
			run_unit<false>(a, C, s, k, cnt, lane);
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
	size_t const smem = static_cast<size_t>(kWarps) * a.tile_cap * sizeof(unsigned short);
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
	long long const want = (units + kWarps - 1) / kWarps;
	int const grid       = static_cast<int>(std::min<long long>(want, static_cast<long long>(sms[device]) * blocks_per_sm[device]));
	deliver_tiles<<<grid, kWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(a);
	return static_cast<int>(cudaGetLastError());
}
}
