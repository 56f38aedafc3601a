// fixed_probability adjacency generation on the GPU, bit-exact with the reference.
//
// Reference: spice::fixed_probability::generate (spice/src/topology.cpp:80-112): ONE sequential
// xoroshiro128+ stream for the whole matrix; per source row, `noise += Exp(1/p - 1)` per draw,
// dst = index + round(noise), the row ends with the first draw whose dst >= dst_count (or when
// index reaches max_degree), and that terminating draw is discarded.  Row r therefore starts at
// stream position sum_{r'<r} (deg_r' + 1): data dependent, which is what makes the loop serial.
//
// B200 design — per chunk of the stream:
//   J  fp_checkpoints : engine states every 1024 positions by polynomial jump (GF(2), spice/util/random.h).
//   A  fp_values      : every stream position in parallel: u -> y = log(u) with the bit-exact glibc restatement
//                       (spice/detail/glibc_log.h), stored as f64; plus exact fixed-point block sums of y.
//   S  fp_segscan     : exclusive scan of the per-segment sums, so that the exact prefix Q(t) = sum_{i<t} q(y_i) of
//                       any position is one block base plus a warp scan of 32 values.
//   N  fp_next        : for EVERY position s of the chunk, where the row that starts at s ends.  With
//                       Z(t) = t K - Q(t + 1) (K = 2^F / scale, an integer) the row from s ends at the first t with
//                       Z(t) - (s K - Q(s)) >= C: Z is increasing, so next(s) = end + 1 is a lower bound in a sorted
//                       array and non-decreasing in s — a warp sweeps 1024 consecutive starts with a sliding window.
//                       C comes in two flavours from interval bounds on round(noise); where they disagree (~1e-8 of
//                       the positions) the entry is marked and decided by an exact replay if the orbit ever lands on it.
//   D  fp_double      : next^(2) .. next^(kHop) by pointer doubling; next is monotone, so the gathers are nearly coalesced.
//   B  fp_chase       : the only sequential part: one thread follows next^(kHop) from the chunk's first row start,
//                       one dependent load per kHop rows.
//   F  fp_fill        : the kHop - 1 row starts between two anchors of the chase, one thread per anchor.
//   C  fp_rows        : element-parallel: entry n of a row is n + round(noise_n), and noise_n is known from the prefix
//                       sums to within delta; where that decides the rounding (all but ~1e-6 of the entries) the entry
//                       is written at once, coalesced.  A row with an undecided entry is replayed by one thread with
//                       the reference's exact float recurrence (fp_rows_exact).  Every row is CHECKED: no entry before
//                       its end may terminate it, its terminating draw must; any failed check is reported
//                       (SPICE_ERR_INTERNAL), never papered over.
//
// Floating-point forms are those of the reference build (g++ 13.3 -O2 -ffast-math
// -march=haswell), taken from its disassembly (DESIGN.md lists them).
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "generator.h"
#include "spice/detail/glibc_log.h"
#include "spice/util/random.h"

namespace spice::gen {
namespace {

constexpr int kSeg        = 1024;        // stream positions generated sequentially by one lane
constexpr int kBlk        = 32;          // positions per prefix-sum block (one warp ballot)
constexpr int kWarpSpan   = 32 * kSeg;   // positions covered by one warp of fp_values
constexpr int kGroupSegs  = 4;           // segments whose checkpoints one fp_checkpoints thread steps through
constexpr int kHopLog     = 5;           // the chase follows next^(2^kHopLog)
constexpr int kHop        = 1 << kHopLog;
constexpr unsigned kAmbig = 0x80000000u; // next[]: the end of the row could not be decided from the bounds
constexpr unsigned kStop  = 0x40000000u; // next^(k)[]: fewer than k rows could be followed inside the chunk
constexpr unsigned kPosMask = 0x3fffffffu;
constexpr double kHalfLo  = 0x1.fffffffffffffp-2; // the reference build's round(): trunc(x + 0.49999999999999994)

#define GEN_CUDA(expr)                                                                          \
	do {                                                                                        \
		cudaError_t e_ = (expr);                                                                \
		if (e_ != cudaSuccess) {                                                                \
			if (err)                                                                            \
				*err = std::string(#expr) + ": " + cudaGetErrorString(e_);                      \
			return 2;                                                                           \
		}                                                                                       \
	} while (0)

struct poly128 {
	unsigned long long lo, hi;
};

struct params {
	long long src, dst;        // matrix shape
	long long col_lo, col_hi;  // kept columns
	long long max_degree;
	double c;                  // 1 - 1/p  (= -scale)
	double inv_fix;            // 2^-F
	double fix;                // 2^F
	double delta;              // bound on |approximate - exact| noise
	long long K;               // round(2^F / scale): one draw's worth of index in units of the fixed-point sums
	long long c_hi, c_lo;      // row end thresholds on Z(t) - Z'(s): may have ended (>= c_hi), has certainly ended (>= c_lo)
};

// ---- device helpers --------------------------------------------------------------------------
__device__ __forceinline__ void xoro_advance(unsigned long long& s0, unsigned long long& s1) {
	unsigned long long const t = s0 ^ s1;
	s0                         = ((s0 << 24) | (s0 >> 40)) ^ t ^ (t << 16);
	s1                         = (t << 37) | (t >> 27);
}

__device__ poly128 mulmod(poly128 a, poly128 b, poly128 P) {
	poly128 acc{0, 0};
	for (int i = 127; i >= 0; i--) {
		bool const carry = acc.hi >> 63;
		acc.hi           = (acc.hi << 1) | (acc.lo >> 63);
		acc.lo <<= 1;
		if (carry) {
			acc.lo ^= P.lo;
			acc.hi ^= P.hi;
		}
		if (((i < 64 ? b.lo >> i : b.hi >> (i - 64)) & 1ull) != 0) {
			acc.lo ^= a.lo;
			acc.hi ^= a.hi;
		}
	}
	return acc;
}

// x86 cvttsd2si (32-bit): out-of-range -> INT_MIN
__device__ __forceinline__ int cvttsd2si32(double x) {
	return (x > -2147483649.0 && x < 2147483648.0) ? __double2int_rz(x) : static_cast<int>(0x80000000u);
}

__device__ __forceinline__ long long quantize(double y, double fix) { return __double2ll_rn(__dmul_rn(y, fix)); }

// ---- kernel J: RNG checkpoints -------------------------------------------------------------------
// ckpt[g * kGroupSegs + k] = engine state at chunk position (g * kGroupSegs + k) * kSeg.
// base_poly = x^(chunk base) mod charpoly; xg[i] = x^(kGroupSegs * kSeg * 2^i) mod charpoly.
struct ckpt_args {
	poly128 charpoly, base_poly;
	poly128 xg[24];
	unsigned long long s0, s1; // stream seed state
	long long groups;
	ulonglong2* ckpt;
};

__global__ void __launch_bounds__(128) fp_checkpoints(ckpt_args a) {
	__shared__ ulonglong2 basis[128]; // T^i s, i < 128
	if (threadIdx.x == 0) {
		unsigned long long s0 = a.s0, s1 = a.s1;
		for (int i = 0; i < 128; i++) {
			basis[i] = make_ulonglong2(s0, s1);
			xoro_advance(s0, s1);
		}
	}
	__syncthreads();
	long long const g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (g >= a.groups)
		return;
	poly128 c = a.base_poly;
	for (int i = 0; i < 24; i++)
		if ((g >> i) & 1)
			c = mulmod(c, a.xg[i], a.charpoly);
	unsigned long long s0 = 0, s1 = 0;
	for (int i = 0; i < 128; i++)
		if (((i < 64 ? c.lo >> i : c.hi >> (i - 64)) & 1ull) != 0) {
			s0 ^= basis[i].x;
			s1 ^= basis[i].y;
		}
	for (int k = 0; k < kGroupSegs; k++) {
		a.ckpt[g * kGroupSegs + k] = make_ulonglong2(s0, s1);
		if (k + 1 < kGroupSegs)
			for (int i = 0; i < kSeg; i++)
				xoro_advance(s0, s1);
	}
}

// ---- kernel A: values --------------------------------------------------------------------------------
// One lane walks one segment of kSeg positions; a warp stages 32x32 tiles in shared memory so the
// f64 stores are full 256-byte lines.
struct values_args {
	ulonglong2 const* ckpt;
	double* y;            // [len]
	long long* blk_local; // [len / kBlk]: exclusive prefix of block sums inside the segment
	long long* seg_sum;   // [len / kSeg]
	long long segs;       // number of segments (multiple of 32)
	double fix;
};

__device__ std::uint64_t g_log_tab[256]; // glibc __log_data.tab, uploaded once per process

__global__ void __launch_bounds__(128) fp_values(values_args a) {
	__shared__ std::uint64_t tab[256];
	__shared__ double tile[4][32][33];
	for (int i = threadIdx.x; i < 256; i += blockDim.x)
		tab[i] = g_log_tab[i];
	__syncthreads();

	int const lane      = threadIdx.x & 31;
	int const warp      = threadIdx.x >> 5;
	long long const wid = static_cast<long long>(blockIdx.x) * 4 + warp;
	long long const seg = wid * 32 + lane;
	if (wid * 32 >= a.segs)
		return;
	ulonglong2 const st   = a.ckpt[seg];
	unsigned long long s0 = st.x, s1 = st.y;
	long long run = 0; // exclusive prefix of block sums within the segment
	for (int b = 0; b < kSeg / kBlk; b++) {
		long long bsum = 0;
#pragma unroll 4
		for (int j = 0; j < kBlk; j++) {
			unsigned long long const r = s0 + s1;
			xoro_advance(s0, s1);
			// generate_canonical<double, true>: ((r >> 11) + 1) * 2^-53 in (0, 1]
			double const u = __dmul_rn(__ull2double_rn((r >> 11) + 1), 0x1p-53);
			double const y = spice::detail::glibc::log_with_table(u, tab);
			tile[warp][lane][j] = y;
			bsum += quantize(y, a.fix);
		}
		a.blk_local[seg * (kSeg / kBlk) + b] = run;
		run += bsum;
		__syncwarp();
		for (int row = 0; row < 32; row++)
			a.y[(wid * 32 + row) * kSeg + b * kBlk + lane] = tile[warp][row][lane];
		__syncwarp();
	}
	a.seg_sum[seg] = run;
}

// ---- kernel S: exclusive scan of segment sums (single block) ---------------------------------------
__global__ void __launch_bounds__(1024) fp_segscan(long long const* seg_sum, long long* seg_base, long long segs) {
	__shared__ long long part[1024];
	long long const per = (segs + 1023) / 1024;
	long long const lo  = threadIdx.x * per;
	long long const hi  = min(segs, lo + per);
	long long sum       = 0;
	for (long long i = lo; i < hi; i++)
		sum += seg_sum[i];
	part[threadIdx.x] = sum;
	__syncthreads();
	for (int off = 1; off < 1024; off <<= 1) { // Hillis-Steele inclusive scan
		long long const v = threadIdx.x >= off ? part[threadIdx.x - off] : 0;
		__syncthreads();
		part[threadIdx.x] += v;
		__syncthreads();
	}
	long long run = part[threadIdx.x] - sum;
	for (long long i = lo; i < hi; i++) {
		seg_base[i] = run;
		run += seg_sum[i];
	}
}

// ---- exact prefixes ---------------------------------------------------------------------------------
struct prefix_view {
	double const* y;
	long long const* blk_local;
	long long const* seg_base;
	long long len; // draws held in y (chunk + overlap)
	double fix;
};

__device__ __forceinline__ long long warp_incl_scan(long long v, int lane) {
#pragma unroll
	for (int off = 1; off < 32; off <<= 1) {
		long long const o = __shfl_up_sync(0xffffffffu, v, off);
		if (lane >= off)
			v += o;
	}
	return v;
}

// Q(t + 1) = sum_{i <= t} q(y_i) for the 32 positions t = 32 blk + lane of one aligned block (chunk frame);
// `own` receives q(y_t).  Positions at or behind `len` contribute nothing.
__device__ __forceinline__ long long block_prefix(prefix_view const& V, long long blk, int lane, long long& own) {
	long long const t = blk * kBlk + lane;
	own               = t < V.len ? quantize(V.y[t], V.fix) : 0;
	long long base    = 0;
	if (blk * kBlk < V.len)
		base = V.seg_base[(blk * kBlk) / kSeg] + V.blk_local[blk];
	return base + warp_incl_scan(own, lane);
}

// Z(t) = t K - Q(t + 1) (and Z'(s) = s K - Q(s)) modulo 2^64: t K alone exceeds 64 bits over a chunk, but only differences
// of nearby positions are ever compared, and those are small
__device__ __forceinline__ long long zval(long long t, long long K, long long prefix) {
	return static_cast<long long>(static_cast<unsigned long long>(t) * static_cast<unsigned long long>(K) - static_cast<unsigned long long>(prefix));
}
__device__ __forceinline__ long long zdiff(long long a, long long b) {
	return static_cast<long long>(static_cast<unsigned long long>(a) - static_cast<unsigned long long>(b));
}

// ---- kernel N: next(s) for every position of the chunk ------------------------------------------------------
struct next_args {
	params P;
	prefix_view V;
	long long ch;       // positions [0, ch) get an entry (multiple of kSeg)
	unsigned* next;     // [ch]: (end of the row that starts at s) + 1, | kAmbig
};

constexpr int kWinBlocks = 8; // window of Z values a warp searches at a time: 256 positions, kept as a ring of blocks

__global__ void __launch_bounds__(128) fp_next(next_args a) {
	__shared__ long long win_s[4][kWinBlocks * kBlk];
	int const lane      = threadIdx.x & 31, warp = threadIdx.x >> 5;
	long long const seg = static_cast<long long>(blockIdx.x) * 4 + warp;
	long long const s0  = seg * kSeg;
	if (s0 >= a.ch)
		return;
	params const& P = a.P;
	long long* win  = win_s[warp];
	long long own;

	// Z'(s0) = s0 K - Q(s0): s0 opens an aligned block, so Q(s0) is that block's base
	long long const q_s0 = block_prefix(a.V, s0 / kBlk, lane, own) - own; // exclusive prefix of this lane's position
	long long const zp0  = zval(s0, P.K, __shfl_sync(0xffffffffu, q_s0, 0));

	// anchor: the first position of a block that certainly does not lie behind the end of the row from s0.
	// 32-ary search over block starts: may_end(t) is monotone in t.
	long long blk_lo = s0 / kBlk;                                   // known: the row has not ended before this block's first draw
	long long span   = (P.max_degree + 2 * kBlk) / kBlk + 1;        // blocks that certainly contain the end
	while (span > 1) {
		long long const st = (span + 31) / 32;
		long long const B  = blk_lo + static_cast<long long>(lane + 1) * st;
		long long const t  = B * kBlk;
		bool may_end       = true;
		if (t < a.V.len) {
			long long const p1 = a.V.seg_base[t / kSeg] + a.V.blk_local[B] + quantize(a.V.y[t], a.V.fix); // Q(t + 1)
			may_end            = (zdiff(zval(t, P.K, p1), zp0) >= P.c_hi) | (t - s0 >= P.max_degree);
		}
		unsigned const m = __ballot_sync(0xffffffffu, may_end);
		int const j      = m ? __ffs(m) - 1 : 32;
		blk_lo += static_cast<long long>(j) * st; // the last probed block start that is not an end
		span = st;
	}
	long long anchor = blk_lo; // first block of the window
	long long filled = blk_lo; // blocks [anchor, filled) of the window hold their Z values (ring slot = block % kWinBlocks)

	for (int i = 0; i < kSeg / kBlk; i++) {
		long long const s   = s0 + i * kBlk + lane;
		long long const p1s = block_prefix(a.V, s / kBlk, lane, own);
		long long const zp  = zval(s, P.K, p1s - own);
		long long const cap = s + P.max_degree; // the draw at which index == max_degree ends the row whatever its value
		long long e         = -1;
		bool certain        = false;
		for (;;) {
			// Z of the window's positions: the ends move on by about a block per batch of starts, so most of the window is
			// still there from the batch before
			if (filled < anchor)
				filled = anchor;
			for (; filled < anchor + kWinBlocks; filled++) {
				long long o2;
				long long const p1 = block_prefix(a.V, filled, lane, o2);
				win[(filled % kWinBlocks) * kBlk + lane] = zval(filled * kBlk + lane, P.K, p1);
			}
			__syncwarp();
			long long const w0 = anchor * kBlk;
			if (e < 0) {
				// first j with Z(w0 + j) - zp >= c_hi (wrap-safe: differences of nearby prefixes are small)
				int lo = 0, hi = kWinBlocks * kBlk; // answer in [lo, hi]; hi = not in this window
				while (lo < hi) {
					int const mid = (lo + hi) >> 1;
					if ((w0 + mid >= a.V.len) | (zdiff(win[(w0 + mid) % (kWinBlocks * kBlk)], zp) >= P.c_hi)) // (nothing is looked for behind the chunk's draws)
						hi = mid;
					else
						lo = mid + 1;
				}
				long long const t = w0 + lo;
				if (lo < kWinBlocks * kBlk && t <= cap) {
					e       = t;
					certain = (zdiff(win[t % (kWinBlocks * kBlk)], zp) >= P.c_lo) | (t == cap);
				} else if (cap < w0 + kWinBlocks * kBlk) {
					e       = cap;
					certain = true;
				}
			}
			__syncwarp();
			if (__all_sync(0xffffffffu, e >= 0))
				break;
			anchor += kWinBlocks; // someone's row runs on: the next window
		}
		a.next[s] = static_cast<unsigned>(e + 1) | (certain ? 0u : kAmbig);
		// rows end in the order they start: the next batch's ends are not before this batch's last one
		anchor = __shfl_sync(0xffffffffu, e, 31) / kBlk;
	}
}

// ---- kernel D: pointer doubling -------------------------------------------------------------------------------
// out[s] = in[in[s]]: twice as many rows ahead.  kStop when that leaves the positions the chunk has entries for, or
// crosses an undecided entry.
__global__ void __launch_bounds__(256) fp_double(unsigned const* in, unsigned* out, long long ch) {
	long long const s = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (s >= ch)
		return;
	unsigned const v = in[s];
	unsigned r       = kStop;
	if (!(v & (kAmbig | kStop)) && static_cast<long long>(v) < ch) {
		unsigned const w = in[v];
		if (!(w & (kAmbig | kStop)))
			r = w;
	}
	out[s] = r;
}

// ---- kernel B: the orbit of row starts --------------------------------------------------------------------------
struct orbit_state {
	long long row;        // next row to start
	long long pos;        // its stream position (global)
	long long exact_rows; // rows whose end was decided by the exact replay
	long long rows_done_in_chunk;
	long long anchors;    // anchors written by this chunk's chase
};

struct chase_args {
	params P;
	double const* y;
	unsigned const* next;  // one row ahead
	unsigned const* hop;   // kHop rows ahead
	long long base;        // global position of the chunk's first draw
	long long ch;
	long long len;         // draws held in y
	long long* row_start;  // [src + 1], global positions
	unsigned* anchor_pos;  // chunk-frame position of every kHop-th row start the chase passed ...
	unsigned* anchor_row;  // ... and its row, relative to the chunk's first row
	int* row_flag;         // [src]: 2 = the row's end came from the exact replay: fp_rows must not judge it by the bounds
	orbit_state* st;
};

__global__ void __launch_bounds__(32) fp_chase(chase_args a) {
	if (threadIdx.x != 0)
		return;
	params const& P      = a.P;
	long long row        = a.st->row;
	long long const row0 = row;
	long long s          = a.st->pos - a.base;
	long long nexact     = a.st->exact_rows;
	long long na         = 0;
	while (row < P.src && s < a.ch) {
		unsigned const h = a.hop[s];
		if (!(h & (kAmbig | kStop)) && row + kHop <= P.src) {
			a.anchor_pos[na] = static_cast<unsigned>(s);
			a.anchor_row[na] = static_cast<unsigned>(row - row0);
			na++;
			s = h;
			row += kHop;
			continue;
		}
		unsigned const v = a.next[s];
		long long nx     = v & kPosMask;
		if (v & kAmbig) { // the bounds could not tell where this row ends: replay it
			double noise = 0;
			int index = 0, dst = 0;
			long long t = s;
			// the recurrence is sequential, its inputs are not: the draws are fetched 16 at a time, one batch ahead (rows
			// of 1e5 draws at 1e6 x 1e6 would otherwise pay a memory round trip per draw); reads behind the row's end stay
			// inside the chunk's overlap margin or are clamped to its last draw
			constexpr int kAhead = 16;
			long long const last = a.len - 1;
			double cur[kAhead], nxt[kAhead];
#pragma unroll
			for (int i = 0; i < kAhead; i++)
				cur[i] = a.y[min(t + i, last)];
			for (bool done = false; !done;) {
#pragma unroll
				for (int i = 0; i < kAhead; i++)
					nxt[i] = a.y[min(t + kAhead + i, last)];
				// only the additions to `noise` depend on each other; the end tests of the batch do not
				double nz[kAhead];
				double run = noise;
#pragma unroll
				for (int i = 0; i < kAhead; i++) {
					run   = __fma_rn(cur[i], P.c, run);
					nz[i] = run;
				}
				int first = kAhead; // the first draw of the batch that ends the row
#pragma unroll
				for (int i = kAhead - 1; i >= 0; i--) {
					int const idx = index + i;
					dst           = idx + cvttsd2si32(__dadd_rn(nz[i], copysign(kHalfLo, nz[i])));
					if ((static_cast<long long>(dst) >= P.dst) | (idx >= P.max_degree))
						first = i;
				}
				done = first < kAhead;
				t += first;
				index += first;
				noise = nz[kAhead - 1];
#pragma unroll
				for (int i = 0; i < kAhead; i++)
					cur[i] = nxt[i];
			}
			nx = t + 1;
			nexact++;
			a.row_flag[row] = 2;
		}
		row++;
		s                = nx;
		a.row_start[row] = a.base + s;
	}
	a.st->row                = row;
	a.st->pos                = a.base + s;
	a.st->exact_rows         = nexact;
	a.st->rows_done_in_chunk = row - row0;
	a.st->anchors            = na;
}

// ---- kernel F: the row starts between two anchors ------------------------------------------------------------------
__global__ void __launch_bounds__(128) fp_fill(unsigned const* next, unsigned const* anchor_pos, unsigned const* anchor_row, long long anchors,
                                               long long row0, long long base, long long* row_start) {
	long long const j = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (j >= anchors)
		return;
	unsigned p         = anchor_pos[j];
	long long const r  = row0 + anchor_row[j];
	for (int i = 1; i <= kHop; i++) {
		p                = next[p] & kPosMask;
		row_start[r + i] = base + p;
	}
}

// ---- kernel C: rows, element-parallel ---------------------------------------------------------------------------------
struct rows_args {
	params P;
	prefix_view V;
	long long base;
	long long const* row_start;
	long long row_lo, row_hi;  // rows of this chunk
	long long* degree;         // [src] kept-column degree (count mode)
	long long* below;          // [src] entries of the row left of col_lo (count mode writes, write mode reads); null: unfiltered
	long long const* offsets;  // [src + 1] output offsets (write mode)
	int* neighbors;
	long long capacity;
	int* row_flag;             // [src] != 0: the row is left to fp_rows_exact
	long long* fix_list;       // rows left to fp_rows_exact by this launch ...
	unsigned* fix_count;       // ... and their number
	int* error;                // bit 0: self-check failed, bit 1: capacity exceeded
	int write;                 // 0 = count kept columns, 1 = write
	int warps_per_row;         // 1 or 4
};

__global__ void __launch_bounds__(128) fp_rows(rows_args a) {
	int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	int const wpr  = a.warps_per_row;
	long long const r = a.row_lo + static_cast<long long>(blockIdx.x) * (4 / wpr) + warp / wpr;
	if (r >= a.row_hi)
		return;
	int const sub = warp % wpr; // this warp's share: aligned blocks sub, sub + wpr, ... of the row
	params const& P = a.P;
	if (a.row_flag[r]) { // decided by the exact replay (now or in an earlier pass): not judged by the bounds
		if (sub == 0 && lane == 0)
			a.fix_list[atomicAdd(a.fix_count, 1u)] = r;
		return;
	}
	long long const s = a.row_start[r] - a.base;
	long long const e = a.row_start[r + 1] - a.base - 1; // the terminating draw
	long long own;
	long long const p1s = block_prefix(a.V, s / kBlk, lane, own);
	long long const qs  = __shfl_sync(0xffffffffu, p1s - own, static_cast<int>(s % kBlk));
	long long const out0 = a.write ? a.offsets[r] - (a.below ? a.below[r] : 0) : 0; // entry n goes to out0 + n
	long long kept = 0, left = 0;
	bool ambiguous = false, bad = false;
	for (long long blk = s / kBlk + sub; blk <= e / kBlk; blk += wpr) {
		long long const p1 = block_prefix(a.V, blk, lane, own);
		long long const t  = blk * kBlk + lane;
		if (t < s || t > e)
			continue;
		long long const n  = t - s;
		double const v     = __dmul_rn(__dmul_rn(__ll2double_rn(p1 - qs), P.inv_fix), P.c);
		long long const rl = __double2ll_rd(v - P.delta + 0.5);
		long long const rh = __double2ll_rd(v + P.delta + 0.5);
		if (t == e) { // must end the row
			bool const ends = (n >= P.max_degree) | (n + rh >= P.dst);
			bad |= !ends;
			continue;
		}
		bad |= (n >= P.max_degree) | (n + rh >= P.dst); // must not
		ambiguous |= rl != rh;
		long long const d = n + rl;
		if (d < P.col_lo)
			left++;
		else if (d < P.col_hi) {
			kept++;
			if (a.write) {
				if (out0 + n < a.capacity)
					a.neighbors[out0 + n] = static_cast<int>(d - P.col_lo);
				else
					bad = true;
			}
		}
	}
	if (__any_sync(0xffffffffu, ambiguous)) {
		// an entry whose rounding the bounds cannot decide: the whole row is replayed exactly
		if (lane == 0 && atomicExch(a.row_flag + r, 1) == 0)
			a.fix_list[atomicAdd(a.fix_count, 1u)] = r;
		return;
	}
	if (__any_sync(0xffffffffu, bad))
		if (lane == 0)
			atomicOr(a.error, 1);
	if (!a.write) {
		for (int off = 16; off; off >>= 1) {
			kept += __shfl_xor_sync(0xffffffffu, kept, off);
			left += __shfl_xor_sync(0xffffffffu, left, off);
		}
		if (lane == 0) {
			if (wpr == 1) {
				a.degree[r] = kept;
				a.below[r]  = left;
			} else {
				atomicAdd(reinterpret_cast<unsigned long long*>(a.degree + r), static_cast<unsigned long long>(kept));
				atomicAdd(reinterpret_cast<unsigned long long*>(a.below + r), static_cast<unsigned long long>(left));
			}
		}
	}
}

// One WARP per listed row replays the reference's exact float recurrence, writes (or counts) the row and checks that
// it ends exactly where the orbit said the next row starts.  The recurrence is sequential (lane 0 runs it, one DFMA of
// latency per draw), its inputs and outputs are not: the warp stages the draws tile by tile in shared memory, one tile
// ahead of the chain, and classifies / stores the tile's targets with all lanes.
constexpr int kExactTile = 256;

__global__ void __launch_bounds__(128) fp_rows_exact(rows_args a) {
	__shared__ double ybuf[4][kExactTile];
	__shared__ int dbuf[4][kExactTile];
	int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned const nfix = *a.fix_count;
	params const& P     = a.P;
	double const* y     = a.V.y;
	for (unsigned f = blockIdx.x * 4 + warp; f < nfix; f += gridDim.x * 4) {
		long long const r  = a.fix_list[f];
		long long const s  = a.row_start[r] - a.base;
		long long const en = a.row_start[r + 1] - a.base - 1; // where the orbit says the terminating draw is
		long long const out0 = a.write ? a.offsets[r] : 0;
		long long kept = 0, left = 0; // warp-uniform totals
		double noise = 0;             // lane 0's chain
		long long t  = s;
		bool bad     = false;
		double reg[kExactTile / 32];
#pragma unroll
		for (int j = 0; j < kExactTile / 32; j++) {
			long long const pos = t + j * 32 + lane;
			reg[j]              = pos < a.V.len ? y[pos] : 0.0;
		}
		for (bool done = false; !done;) {
#pragma unroll
			for (int j = 0; j < kExactTile / 32; j++)
				ybuf[warp][j * 32 + lane] = reg[j];
			__syncwarp();
#pragma unroll
			for (int j = 0; j < kExactTile / 32; j++) { // the next tile's draws are on their way while the chain runs
				long long const pos = t + kExactTile + j * 32 + lane;
				reg[j]              = pos < a.V.len ? y[pos] : 0.0;
			}
			// lane 0 runs nothing but the chain: noise_i = fma(y_i, c, noise_(i-1)), written over the draw it consumed
			if (lane == 0) {
				double run = noise;
#pragma unroll 8
				for (int i = 0; i < kExactTile; i++) {
					run            = __fma_rn(ybuf[warp][i], P.c, run);
					ybuf[warp][i]  = run;
				}
				noise = run;
			}
			__syncwarp();
			// the roundings, the targets and the end tests of the tile's draws are independent: all lanes
			int cnt = kExactTile, state = 0; // state: 1 = the row ended inside this tile, 2 = it did not end where the orbit says
			for (int i0 = 0; i0 < kExactTile && state == 0; i0 += 32) {
				int const i        = i0 + lane;
				long long const ix = (t - s) + i; // the entry's index in its row
				double const nz    = ybuf[warp][i];
				int const d        = static_cast<int>(ix) + cvttsd2si32(__dadd_rn(nz, copysign(kHalfLo, nz)));
				bool const past    = t + i > en; // would run past the orbit's row end: reported below
				bool const stop    = (static_cast<long long>(d) >= P.dst) | (ix >= P.max_degree);
				unsigned const m   = __ballot_sync(0xffffffffu, past | stop);
				if (m) {
					int const first = __ffs(m) - 1;
					bool const p1   = __shfl_sync(0xffffffffu, past, first);
					cnt             = i0 + first;
					state           = (!p1 && t + cnt == en) ? 1 : 2;
				}
				if (i < cnt)
					dbuf[warp][i] = d;
			}
			__syncwarp();
			// the tile's entries: targets ascend along a row, so the ones left of col_lo come first, then the kept ones
			for (int i0 = 0; i0 < cnt; i0 += 32) {
				int const i        = i0 + lane;
				int const d        = i < cnt ? dbuf[warp][i] : 0x7fffffff;
				bool const is_left = i < cnt && d < P.col_lo;
				bool const is_kept = i < cnt && d >= P.col_lo && d < P.col_hi;
				unsigned const ml  = __ballot_sync(0xffffffffu, is_left);
				unsigned const mk  = __ballot_sync(0xffffffffu, is_kept);
				left += __popc(ml);
				if (a.write && is_kept) {
					long long const o = out0 + kept + __popc(mk & ((1u << lane) - 1));
					if (o < a.capacity)
						a.neighbors[o] = d - static_cast<int>(P.col_lo);
					else
						bad = true;
				}
				kept += __popc(mk);
			}
			__syncwarp();
			t += cnt;
			bad |= state == 2;
			done = state != 0;
		}
		if (__any_sync(0xffffffffu, bad) && lane == 0)
			atomicOr(a.error, bad && a.write && out0 + kept > a.capacity ? 3 : 1);
		if (lane == 0 && !a.write) {
			a.degree[r] = kept;
			if (a.below)
				a.below[r] = left;
		}
	}
}

// offsets = exclusive scan(degree) when columns are filtered; closed form otherwise
__global__ void fp_offsets_closed_form(long long const* row_start, long long* offsets, long long lo, long long hi) {
	long long const r = lo + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (r <= hi)
		offsets[r] = row_start[r] - r; // draws before row r minus one discarded draw per earlier row
}

__global__ void __launch_bounds__(1024) fp_scan_degrees(long long const* degree, long long* offsets, long long src) {
	__shared__ long long part[1024];
	long long const per = (src + 1023) / 1024;
	long long const lo  = threadIdx.x * per;
	long long const hi  = min(src, lo + per);
	long long sum       = 0;
	for (long long i = lo; i < hi; i++)
		sum += degree[i];
	part[threadIdx.x] = sum;
	__syncthreads();
	for (int off = 1; off < 1024; off <<= 1) {
		long long const v = threadIdx.x >= off ? part[threadIdx.x - off] : 0;
		__syncthreads();
		part[threadIdx.x] += v;
		__syncthreads();
	}
	long long run = part[threadIdx.x] - sum;
	for (long long i = lo; i < hi; i++) {
		offsets[i] = run;
		run += degree[i];
	}
	if (threadIdx.x == 1023)
		offsets[src] = part[1023];
}

poly128 to_dev(util::jump::poly p) { return {p.lo, p.hi}; }

long long round_up(long long v, long long m) { return (v + m - 1) / m * m; }
} // namespace

long long max_degree(long long dst, double p) {
	double const dp = static_cast<double>(dst) * p;
	return static_cast<long long>(std::fma(std::sqrt((1.0 - p) * dp), 3.0, dp)); // topology.cpp:75-78 as compiled
}

namespace {
// p == 1: scale == 0, noise stays 0, every row is 0 .. min(dst, max_degree) - 1 and consumes one draw more than that
__global__ void fp_dense_rows(long long src, long long row_len, long long col_lo, long long col_hi, long long* offsets, int* neighbors) {
	long long const lo = min(col_lo, row_len), hi = min(col_hi, row_len);
	long long const w  = hi - lo;
	long long const i  = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (i <= src)
		offsets[i] = i * w;
	if (i < src * w)
		neighbors[i] = static_cast<int>(i % w + lo - col_lo);
}

// everything a generation allocates; released on every path out of generate_fixed_probability
struct scratch {
	std::vector<void*> dev;
	std::vector<void*> host;
	std::vector<cudaEvent_t> events;
	template <class T>
	cudaError_t alloc(T** p, size_t n) {
		cudaError_t const e = cudaMalloc(reinterpret_cast<void**>(p), std::max<size_t>(n * sizeof(T), 16));
		if (e == cudaSuccess)
			dev.push_back(*p);
		return e;
	}
	cudaError_t event(cudaEvent_t* e) {
		cudaError_t const r = cudaEventCreate(e);
		if (r == cudaSuccess)
			events.push_back(*e);
		return r;
	}
	~scratch() {
		for (void* p : dev)
			cudaFree(p);
		for (void* p : host)
			cudaFreeHost(p);
		for (cudaEvent_t e : events)
			cudaEventDestroy(e);
	}
};
}

int generate_fixed_probability(void* stream_, long long src, long long dst, double p, unsigned long long seed_lo,
                               unsigned long long seed_hi, long long col_lo, long long col_hi, long long chunk_draws,
                               result* out, std::string* err) {
	auto stream = static_cast<cudaStream_t>(stream_);
	*out        = result{};
	// SPICE_GEN_TIMING=1: host wall-clock of the phases on stderr (synchronises at every lap)
	bool const timing = std::getenv("SPICE_GEN_TIMING") != nullptr;
	auto t_last       = std::chrono::steady_clock::now();
	auto lap          = [&](char const* what) {
        if (!timing)
            return;
        cudaStreamSynchronize(stream);
        auto const now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[spice gen %lld x %lld] %-24s %8.2f ms\n", src, dst, what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
	};
	scratch S; // frees the scratch buffers and events on every return below; the result's arrays belong to the caller
	cudaEvent_t ev0 = nullptr, ev1 = nullptr, evr0 = nullptr, evr1 = nullptr;
	GEN_CUDA(S.event(&ev0));
	GEN_CUDA(S.event(&ev1));
	GEN_CUDA(S.event(&evr0));
	GEN_CUDA(S.event(&evr1));
	GEN_CUDA(cudaEventRecord(ev0, stream));

	GEN_CUDA(cudaMalloc(&out->offsets, sizeof(long long) * static_cast<size_t>(src + 1)));
	GEN_CUDA(cudaMemsetAsync(out->offsets, 0, sizeof(long long) * static_cast<size_t>(src + 1), stream));
	if (src == 0 || dst == 0 || p == 0 || col_hi <= col_lo) { // topology.cpp:85-86
		GEN_CUDA(cudaStreamSynchronize(stream));
		return 0;
	}

	params P{};
	P.src        = src;
	P.dst        = dst;
	P.col_lo     = col_lo;
	P.col_hi     = col_hi;
	P.max_degree = max_degree(dst, p);
	P.c          = 1.0 - 1.0 / p;
	double const scale = -P.c;
	if (scale * 37.0 + static_cast<double>(dst) >= 2147483000.0) {
		if (err)
			*err = "fixed_probability: p too small for 32-bit target arithmetic";
		return 3;
	}
	if (scale == 0) { // p == 1
		long long const row_len = std::min(dst, P.max_degree);
		long long const w       = std::min(col_hi, row_len) - std::min(col_lo, row_len);
		GEN_CUDA(cudaMalloc(&out->neighbors, sizeof(int) * static_cast<size_t>(std::max<long long>(src * w, 0) + 8)));
		long long const n = std::max(src * w, src + 1);
		fp_dense_rows<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(src, row_len, col_lo, col_hi, out->offsets, out->neighbors);
		GEN_CUDA(cudaGetLastError());
		GEN_CUDA(cudaEventRecord(ev1, stream));
		GEN_CUDA(cudaStreamSynchronize(stream));
		out->edges    = src * w;
		out->draws    = src * (row_len + 1);
		out->launches = 1;
		cudaEventElapsedTime(&out->total_ms, ev0, ev1);
		return 0;
	}
	// fixed-point scale 2^F of the prefix sums: |sum of y over one row| <= (max_degree + 1) * 36.8 must stay below 2^60,
	// and so must dst * K, K = 2^F / scale (the row-end thresholds below)
	int bits = 1;
	while (std::ldexp(1.0, bits) < (static_cast<double>(P.max_degree) + 2.0) * 37.0)
		bits++;
	int F = std::min(44, 60 - bits);
	while (F > 8 && (static_cast<double>(dst) + 2.0) * std::ldexp(1.0, F) / scale >= std::ldexp(1.0, 60))
		F--;
	P.fix       = std::ldexp(1.0, F);
	P.inv_fix   = std::ldexp(1.0, -F);
	double const nmax = static_cast<double>(P.max_degree) + 2.0;
	// |float-sequential noise - exact sum|: <= n * ulp(dst)/2 ; quantisation: n * 2^-(F+1) * scale ; slack x4
	P.delta = 4.0 * (nmax * (std::ldexp(static_cast<double>(dst) + scale * 37.0 + 64.0, -52) + std::ldexp(scale + 1.0, -(F + 1)))) + 1e-9;
	if (!(P.delta < 0.125)) {
		if (err)
			*err = "fixed_probability: p too close to 1 for the prefix-sum bounds";
		return 3;
	}
	{
		// n + round(noise) >= dst  <=>  noise >= dst - n - 0.5 (up to delta)  <=>  Z(t) - Z'(s) >= (dst - 0.5 -+ delta) K,
		// evaluated with the integer K: the difference is at most n / 2 <= max_degree units, added as slack
		long double const Kr = static_cast<long double>(P.fix) / static_cast<long double>(scale);
		P.K    = static_cast<long long>(llroundl(Kr));
		P.c_hi = static_cast<long long>(floorl((static_cast<long double>(dst) - 0.5L - static_cast<long double>(P.delta)) * Kr)) - P.max_degree - 4;
		P.c_lo = static_cast<long long>(ceill((static_cast<long double>(dst) - 0.5L + static_cast<long double>(P.delta)) * Kr)) + P.max_degree + 4;
	}

	// SPICE_GEN_FORCE_EXACT=1 (tests): pretend the bounds decide nothing — every row end is found by the chase's exact
	// replay and every row is written by fp_rows_exact, the paths that otherwise see one row in 1e8 / one in 100
	if (std::getenv("SPICE_GEN_FORCE_EXACT")) {
		P.delta = 0.1249;
		P.c_lo  = (1ll << 62);
	}

	// expected stream length and output size
	double const mean_deg = std::min(static_cast<double>(dst) * p + 1.0, static_cast<double>(P.max_degree));
	long long const ov    = round_up(P.max_degree + 2 + (kWinBlocks + 2) * kBlk, kSeg);
	long long ch          = chunk_draws > 0 ? chunk_draws : (1ll << 26);
	{
		double const est = static_cast<double>(src) * (mean_deg + 1.0) * 1.02 + 65536.0;
		if (est < static_cast<double>(ch))
			ch = static_cast<long long>(est);
	}
	ch                  = round_up(std::max<long long>(ch, kWarpSpan), kWarpSpan);
	long long const len = round_up(ch + ov, kWarpSpan);
	if (len >= (1ll << 30)) {
		if (err)
			*err = "fixed_probability: chunk too long for 30-bit positions";
		return 3;
	}
	long long const segs = len / kSeg;
	long long const groups = (segs + kGroupSegs - 1) / kGroupSegs;

	double const kept_frac = static_cast<double>(col_hi - col_lo) / static_cast<double>(dst);
	double const exp_edges = static_cast<double>(src) * static_cast<double>(dst) * p * kept_frac;
	long long capacity     = static_cast<long long>(exp_edges + 8.0 * std::sqrt(exp_edges + 1.0) + 4096.0);
	capacity               = std::min(capacity, src * std::min(P.max_degree, col_hi - col_lo));
	capacity               = std::max<long long>(capacity, 1);

	double* y            = nullptr;
	long long *blk_local = nullptr, *seg_sum = nullptr, *seg_base = nullptr, *row_start = nullptr, *degree = nullptr, *below = nullptr,
	          *fix_list  = nullptr;
	unsigned *next = nullptr, *hop_a = nullptr, *hop_b = nullptr, *anchor_pos = nullptr, *anchor_row = nullptr, *fix_count = nullptr;
	int* row_flag        = nullptr;
	ulonglong2* ckpt     = nullptr;
	orbit_state* st      = nullptr;
	int* error           = nullptr;
	orbit_state* st_host = nullptr;
	bool const filtered  = !(col_lo == 0 && col_hi == dst);
	long long const max_anchors = ch / kHop + 2;
	GEN_CUDA(S.alloc(&y, static_cast<size_t>(len)));
	GEN_CUDA(S.alloc(&blk_local, static_cast<size_t>(len / kBlk)));
	GEN_CUDA(S.alloc(&seg_sum, static_cast<size_t>(segs)));
	GEN_CUDA(S.alloc(&seg_base, static_cast<size_t>(segs)));
	GEN_CUDA(S.alloc(&ckpt, static_cast<size_t>(groups * kGroupSegs)));
	GEN_CUDA(S.alloc(&row_start, static_cast<size_t>(src + 1)));
	GEN_CUDA(S.alloc(&next, static_cast<size_t>(ch)));
	GEN_CUDA(S.alloc(&hop_a, static_cast<size_t>(ch)));
	GEN_CUDA(S.alloc(&hop_b, static_cast<size_t>(ch)));
	GEN_CUDA(S.alloc(&anchor_pos, static_cast<size_t>(max_anchors)));
	GEN_CUDA(S.alloc(&anchor_row, static_cast<size_t>(max_anchors)));
	GEN_CUDA(S.alloc(&row_flag, static_cast<size_t>(src)));
	GEN_CUDA(S.alloc(&fix_list, static_cast<size_t>(2 * src))); // a row is listed at most twice (by two of its warps)
	GEN_CUDA(S.alloc(&fix_count, 1));
	GEN_CUDA(S.alloc(&st, 1));
	GEN_CUDA(S.alloc(&error, 1));
	GEN_CUDA(cudaMallocHost(&st_host, sizeof(orbit_state)));
	S.host.push_back(st_host);
	GEN_CUDA(cudaMalloc(&out->neighbors, sizeof(int) * static_cast<size_t>(capacity + 8))); // +8: the delivery kernel reads whole 16-byte groups
	if (filtered) {
		GEN_CUDA(S.alloc(&degree, static_cast<size_t>(src)));
		GEN_CUDA(S.alloc(&below, static_cast<size_t>(src)));
		GEN_CUDA(cudaMemsetAsync(degree, 0, sizeof(long long) * static_cast<size_t>(src), stream));
		GEN_CUDA(cudaMemsetAsync(below, 0, sizeof(long long) * static_cast<size_t>(src), stream));
	}
	GEN_CUDA(cudaMemsetAsync(st, 0, sizeof(orbit_state), stream));
	GEN_CUDA(cudaMemsetAsync(error, 0, sizeof(int), stream));
	GEN_CUDA(cudaMemsetAsync(row_flag, 0, sizeof(int) * static_cast<size_t>(src), stream));
	GEN_CUDA(cudaMemsetAsync(row_start, 0, sizeof(long long), stream)); // row 0 starts at position 0
	lap("allocations");

	{
		static bool uploaded[64] = {};
		static std::mutex upload_mutex; // generations of one network run concurrently (runtime.cu build_pending)
		std::lock_guard<std::mutex> lock(upload_mutex);
		int dev = 0;
		GEN_CUDA(cudaGetDevice(&dev));
		if (dev >= 0 && dev < 64 && !uploaded[dev]) {
			GEN_CUDA(cudaMemcpyToSymbol(g_log_tab, spice::detail::glibc::log_tab, sizeof(g_log_tab)));
			uploaded[dev] = true;
		}
	}
	ckpt_args ca{};
	ca.charpoly = to_dev(util::jump::charpoly());
	{
		util::jump::poly g = util::jump::xpow(static_cast<UInt>(kGroupSegs) * kSeg);
		for (int i = 0; i < 24; i++) {
			ca.xg[i] = to_dev(g);
			g        = util::jump::mulmod(g, g);
		}
	}
	ca.s0     = seed_lo;
	ca.s1     = seed_hi;
	ca.groups = groups;
	ca.ckpt   = ckpt;

	struct chunk_rows {
		long long base, row_lo, row_hi;
	};
	std::vector<chunk_rows> chunks; // only used in filtered mode (second pass)
	float rows_ms = 0;
	long long row = 0, base = 0;
	int launches = 0;
	prefix_view const V{y, blk_local, seg_base, len, P.fix};
	// rows of ~100 entries and more: four warps to a row; shorter ones: one
	int const warps_per_row = mean_deg >= 256 ? 4 : 1;

	auto run_values = [&](long long chunk_base) -> int {
		ca.base_poly = to_dev(util::jump::xpow(static_cast<UInt>(chunk_base)));
		fp_checkpoints<<<static_cast<int>((groups + 127) / 128), 128, 0, stream>>>(ca);
		values_args va{ckpt, y, blk_local, seg_sum, segs, P.fix};
		fp_values<<<static_cast<int>((segs / 32 + 3) / 4), 128, 0, stream>>>(va);
		fp_segscan<<<1, 1024, 0, stream>>>(seg_sum, seg_base, segs);
		launches += 3;
		return static_cast<int>(cudaGetLastError());
	};
	auto run_rows = [&](long long chunk_base, long long rlo, long long rhi, int write) -> int {
		if (rhi <= rlo)
			return 0;
		rows_args ra{P,        V,        chunk_base, row_start, rlo,   rhi,   degree,       filtered ? below : nullptr, out->offsets, out->neighbors,
		             capacity, row_flag, fix_list,   fix_count, error, write, warps_per_row};
		cudaMemsetAsync(fix_count, 0, sizeof(unsigned), stream);
		cudaEventRecord(evr0, stream);
		int const rows_per_cta = 4 / warps_per_row;
		fp_rows<<<static_cast<unsigned>((rhi - rlo + rows_per_cta - 1) / rows_per_cta), 128, 0, stream>>>(ra);
		fp_rows_exact<<<296, 128, 0, stream>>>(ra); // rows with an entry (or an end) the bounds could not decide: strided over the list
		cudaEventRecord(evr1, stream);
		launches += 2;
		return static_cast<int>(cudaGetLastError());
	};

	lap("jump polynomials");
	while (row < src) {
		GEN_CUDA(static_cast<cudaError_t>(run_values(base)));
		lap("values");
		next_args na{P, V, ch, next};
		fp_next<<<static_cast<unsigned>((ch / kSeg + 3) / 4), 128, 0, stream>>>(na);
		unsigned const dgrid = static_cast<unsigned>((ch + 255) / 256);
		unsigned const* from = next;
		for (int k = 0; k < kHopLog; k++) {
			unsigned* to = (k & 1) ? hop_b : hop_a;
			fp_double<<<dgrid, 256, 0, stream>>>(from, to, ch);
			from = to;
		}
		chase_args cha{P, y, next, from, base, ch, len, row_start, anchor_pos, anchor_row, row_flag, st};
		fp_chase<<<1, 32, 0, stream>>>(cha);
		launches += 2 + kHopLog;
		GEN_CUDA(cudaGetLastError());
		GEN_CUDA(cudaMemcpyAsync(st_host, st, sizeof(orbit_state), cudaMemcpyDeviceToHost, stream));
		GEN_CUDA(cudaStreamSynchronize(stream));
		lap("next + doubling + chase");
		long long const rhi = st_host->row;
		if (st_host->anchors > 0) {
			fp_fill<<<static_cast<unsigned>((st_host->anchors + 127) / 128), 128, 0, stream>>>(next, anchor_pos, anchor_row, st_host->anchors, row, base,
			                                                                                   row_start);
			launches++;
		}
		if (!filtered) {
			// offsets are a closed form of the row starts, so rows can be written right away
			fp_offsets_closed_form<<<static_cast<int>((rhi - row + 1 + 255) / 256), 256, 0, stream>>>(row_start, out->offsets, row, rhi);
			launches++;
			GEN_CUDA(static_cast<cudaError_t>(run_rows(base, row, rhi, 1)));
			GEN_CUDA(cudaStreamSynchronize(stream));
			float ms = 0;
			if (rhi > row && cudaEventElapsedTime(&ms, evr0, evr1) == cudaSuccess)
				rows_ms += ms;
		} else {
			GEN_CUDA(static_cast<cudaError_t>(run_rows(base, row, rhi, 0)));
			chunks.push_back({base, row, rhi});
		}
		lap("rows");
		if (rhi == row && st_host->pos < base + ch) {
			if (err)
				*err = "fixed_probability: orbit made no progress";
			return 4;
		}
		row  = rhi;
		base = base + ch;
	}
	out->draws      = st_host->pos;
	out->exact_rows = st_host->exact_rows;

	if (filtered) {
		fp_scan_degrees<<<1, 1024, 0, stream>>>(degree, out->offsets, src);
		launches++;
		for (auto const& cr : chunks) {
			if (chunks.size() > 1) // with a single chunk the values are still in the buffer
				GEN_CUDA(static_cast<cudaError_t>(run_values(cr.base)));
			GEN_CUDA(static_cast<cudaError_t>(run_rows(cr.base, cr.row_lo, cr.row_hi, 1)));
			GEN_CUDA(cudaStreamSynchronize(stream));
			float ms = 0;
			if (cr.row_hi > cr.row_lo && cudaEventElapsedTime(&ms, evr0, evr1) == cudaSuccess)
				rows_ms += ms;
		}
	}

	long long edges = 0;
	int herr        = 0;
	GEN_CUDA(cudaMemcpyAsync(&edges, out->offsets + src, sizeof(long long), cudaMemcpyDeviceToHost, stream));
	GEN_CUDA(cudaMemcpyAsync(&herr, error, sizeof(int), cudaMemcpyDeviceToHost, stream));
	GEN_CUDA(cudaEventRecord(ev1, stream));
	GEN_CUDA(cudaStreamSynchronize(stream));
	lap("tail");
	out->edges    = edges;
	out->launches = launches;
	out->rows_ms  = rows_ms;
	cudaEventElapsedTime(&out->total_ms, ev0, ev1);

	if (herr & 1) {
		if (err)
			*err = "fixed_probability: self-check failed (a row did not end where the orbit predicted)";
		return 4;
	}
	if (herr & 2) {
		if (err)
			*err = "fixed_probability: neighbor capacity exceeded";
		return 4;
	}
	return 0;
}

// ---- adj_list ---------------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) adj_pack(int const* es, int const* ed, long long n, long long src, long long dst,
                                                unsigned long long* keys, int* flags) {
	long long const i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	int const s = es[i], d = ed[i];
	if (s < 0 || s >= src || d < 0 || d >= dst) {
		atomicOr(flags, 1); // out of range
		keys[i] = ~0ull;
		return;
	}
	keys[i] = (static_cast<unsigned long long>(static_cast<unsigned>(s)) << 32) | static_cast<unsigned>(d);
}
// sorted keys -> per-row counts of kept targets, multapse flag
__global__ void __launch_bounds__(256) adj_count(unsigned long long const* keys, long long n, long long col_lo, long long col_hi,
                                                 unsigned long long* degree, int* flags) {
	long long const i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	unsigned long long const k = keys[i];
	if (i > 0 && keys[i - 1] == k)
		atomicOr(flags, 2);
	long long const d = static_cast<long long>(k & 0xffffffffull);
	if (d >= col_lo && d < col_hi)
		atomicAdd(degree + (k >> 32), 1ull);
}
// kept targets in key order: entry i goes to offsets[src] + (its rank among the kept keys of its row); keys are sorted,
// so that rank is (number of kept keys before i) - offsets[src] = kept_before[i] - offsets[src]
__global__ void __launch_bounds__(256) adj_write(unsigned long long const* keys, long long const* kept_before, long long n,
                                                 long long col_lo, long long col_hi, int* neighbors) {
	long long const i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	long long const d = static_cast<long long>(keys[i] & 0xffffffffull);
	if (d >= col_lo && d < col_hi)
		neighbors[kept_before[i]] = static_cast<int>(d - col_lo);
}
__global__ void __launch_bounds__(256) adj_keep_flags(unsigned long long const* keys, long long n, long long col_lo, long long col_hi,
                                                      long long* keep) {
	long long const i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	long long const d = static_cast<long long>(keys[i] & 0xffffffffull);
	keep[i]           = (d >= col_lo && d < col_hi) ? 1 : 0;
}
}

int generate_adj_list(void* stream_, int const* edges_src, int const* edges_dst, long long n, long long src, long long dst, long long col_lo,
                      long long col_hi, result* out, bool* duplicates, std::string* err) {
	auto stream = static_cast<cudaStream_t>(stream_);
	*out        = result{};
	if (duplicates)
		*duplicates = false;
	scratch S;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	GEN_CUDA(S.event(&ev0));
	GEN_CUDA(S.event(&ev1));
	GEN_CUDA(cudaEventRecord(ev0, stream));
	GEN_CUDA(cudaMalloc(&out->offsets, sizeof(long long) * static_cast<size_t>(src + 1)));
	GEN_CUDA(cudaMemsetAsync(out->offsets, 0, sizeof(long long) * static_cast<size_t>(src + 1), stream));
	if (n <= 0 || src <= 0) {
		GEN_CUDA(cudaMalloc(&out->neighbors, sizeof(int) * 8));
		GEN_CUDA(cudaStreamSynchronize(stream));
		return 0;
	}
	int *es = nullptr, *ed = nullptr, *flags = nullptr;
	unsigned long long *keys = nullptr, *keys2 = nullptr, *degree = nullptr;
	long long* kept = nullptr;
	void* tmp       = nullptr;
	GEN_CUDA(S.alloc(&es, static_cast<size_t>(n)));
	GEN_CUDA(S.alloc(&ed, static_cast<size_t>(n)));
	GEN_CUDA(S.alloc(&keys, static_cast<size_t>(n)));
	GEN_CUDA(S.alloc(&keys2, static_cast<size_t>(n)));
	GEN_CUDA(S.alloc(&kept, static_cast<size_t>(n)));
	GEN_CUDA(S.alloc(&degree, static_cast<size_t>(src + 1)));
	GEN_CUDA(S.alloc(&flags, 1));
	GEN_CUDA(cudaMemsetAsync(flags, 0, sizeof(int), stream));
	GEN_CUDA(cudaMemsetAsync(degree, 0, sizeof(unsigned long long) * static_cast<size_t>(src + 1), stream));
	GEN_CUDA(cudaMemcpyAsync(es, edges_src, sizeof(int) * static_cast<size_t>(n), cudaMemcpyHostToDevice, stream));
	GEN_CUDA(cudaMemcpyAsync(ed, edges_dst, sizeof(int) * static_cast<size_t>(n), cudaMemcpyHostToDevice, stream));
	unsigned const grid = static_cast<unsigned>((n + 255) / 256);
	adj_pack<<<grid, 256, 0, stream>>>(es, ed, n, src, dst, keys, flags);
	// the key's significant bits: 32 of the target, those of the largest source above them
	int src_bits = 1;
	while ((1ll << src_bits) < src)
		src_bits++;
	int dst_bits = 1;
	while ((1ll << dst_bits) < dst)
		dst_bits++;
	size_t bytes = 0, bytes2 = 0;
	cub::DoubleBuffer<unsigned long long> db(keys, keys2);
	GEN_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, bytes, db, static_cast<long long>(n), 0, 32 + src_bits, stream));
	GEN_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes2, kept, kept, static_cast<long long>(n), stream));
	bytes = std::max(bytes, bytes2);
	GEN_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes2, reinterpret_cast<long long*>(degree), out->offsets, static_cast<long long>(src + 1), stream));
	bytes = std::max(bytes, bytes2);
	GEN_CUDA(cudaMalloc(&tmp, std::max<size_t>(bytes, 16)));
	S.dev.push_back(tmp);
	if (dst_bits < 32) { // two passes: the low 32 bits hold a target below 2^dst_bits, the bits in between are zero
		GEN_CUDA(cub::DeviceRadixSort::SortKeys(tmp, bytes, db, static_cast<long long>(n), 0, dst_bits, stream));
		GEN_CUDA(cub::DeviceRadixSort::SortKeys(tmp, bytes, db, static_cast<long long>(n), 32, 32 + src_bits, stream));
	} else
		GEN_CUDA(cub::DeviceRadixSort::SortKeys(tmp, bytes, db, static_cast<long long>(n), 0, 32 + src_bits, stream));
	unsigned long long const* sorted = db.Current();
	adj_count<<<grid, 256, 0, stream>>>(sorted, n, col_lo, col_hi, degree, flags);
	adj_keep_flags<<<grid, 256, 0, stream>>>(sorted, n, col_lo, col_hi, kept);
	GEN_CUDA(cub::DeviceScan::ExclusiveSum(tmp, bytes, kept, kept, static_cast<long long>(n), stream));
	GEN_CUDA(cub::DeviceScan::ExclusiveSum(tmp, bytes, reinterpret_cast<long long*>(degree), out->offsets, static_cast<long long>(src + 1), stream));
	long long edges = 0;
	int hflags      = 0;
	GEN_CUDA(cudaMemcpyAsync(&edges, out->offsets + src, sizeof(long long), cudaMemcpyDeviceToHost, stream));
	GEN_CUDA(cudaMemcpyAsync(&hflags, flags, sizeof(int), cudaMemcpyDeviceToHost, stream));
	GEN_CUDA(cudaStreamSynchronize(stream));
	if (hflags & 1) {
		if (err)
			*err = "adj_list: a source or target index is out of range";
		return 1;
	}
	GEN_CUDA(cudaMalloc(&out->neighbors, sizeof(int) * static_cast<size_t>(edges + 8))); // +8: the delivery kernel reads whole 16-byte groups
	adj_write<<<grid, 256, 0, stream>>>(sorted, kept, n, col_lo, col_hi, out->neighbors);
	GEN_CUDA(cudaGetLastError());
	GEN_CUDA(cudaEventRecord(ev1, stream));
	GEN_CUDA(cudaStreamSynchronize(stream));
	out->edges    = edges;
	out->launches = 12;
	cudaEventElapsedTime(&out->total_ms, ev0, ev1);
	if (duplicates)
		*duplicates = (hflags & 2) != 0;
	return 0;
}

// ---- fixed_probability, counter-based ------------------------------------------------------------------------------------
// The generator the north star describes ("counter-based RNG and geometric skip sampling ... bounded by write bandwidth"),
// NOT the reference's stream: independent Bernoulli(p) per (source, target) pair, drawn as geometric gaps between connected
// targets by engines addressed with seed_seq::stream(id) (random.h:169, which the reference defines and never uses).  Same
// distribution family as the reference's sampler (which rounds an exponential and truncates rows at mean + 3 sigma), not
// the same matrix.  Two definitions, by p (both restated in tests/test_gpu_generator.py):
//
//  p >= 2^-13  "tiles" (fp_fast_tiles, ONE pass):  a row is cut into blocks of B = 2^ceil(log2(512 / p)) consecutive
//     targets (geometric gaps are memoryless, so restarting at block starts changes nothing), unit (row r, block b) has the
//     engines stream((r * nblocks + b) * 32 + lane).  Per iteration every lane takes two 64-bit draws = four 32-bit
//     uniforms (high word first), each mapped to gap = 1 + #{k >= 1 : u < T[k]}, T[k] = floor(2^32 (1-p)^k) — an integer
//     table built on the host, so the map is exact and restatable; the device finds k from a MUFU.LG2 estimate and fixes it
//     against the table.  Lane l's four gaps follow lane l-1's; a warp scan of the lane sums turns them into targets.  A
//     warp collects its unit in shared memory; a CTA of 8 warps = one tile of 8 consecutive units, tiles are claimed in
//     order from a ticket and get their place in the output by decoupled look-back over (status, count) words, so the
//     neighbours are written once, coalesced, with no counting pass and no host round trip (the output is allocated at
//     mean + 10 sigma entries; the run repeats with the exact size in the never-seen case that this is too small).
//     A rank generates only the blocks that overlap its columns.  fp_fast_tiles_piped (the default) takes the look-back
//     off the generators' path: two buffers per generator warp, a ninth warp claims the tickets and places the tiles.
//  p <  2^-13  "log path" (fp_fast_rows, count + write passes): a warp per row, engines stream(r * 32 + lane),
//     gap = 1 + floor(log(u) / log(1 - p)) in double with the glibc log restatement, 32 gaps per iteration.
namespace {
struct fast_args {
	long long src, dst, col_lo, col_hi;
	unsigned long long seed_lo, seed_hi;
	double inv_log1mp;         // 1 / log(1 - p)   (-0.0 for p == 1: every gap is 1)
	long long* degree;         // [src + 1] kept targets per row (count pass writes)
	long long const* offsets;  // [src + 1] (write pass reads)
	int* neighbors;
};

template <bool kWrite>
__global__ void __launch_bounds__(256) fp_fast_rows(fast_args a) {
	__shared__ std::uint64_t tab[256];
	for (int i = threadIdx.x; i < 256; i += blockDim.x)
		tab[i] = g_log_tab[i];
	__syncthreads();
	int const lane       = threadIdx.x & 31;
	long long const w    = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
	long long const W    = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
	util::seed_seq const seed(UInt128{a.seed_lo, a.seed_hi});
	for (long long r = w; r < a.src; r += W) {
		UInt128 const st = seed.stream(static_cast<UInt>(r) * 32 + static_cast<UInt>(lane)).seed();
		unsigned long long s0 = st.lo, s1 = st.hi;
		long long base = -1; // the last target taken so far
		long long kept = 0;
		long long const out0 = kWrite ? a.offsets[r] : 0;
		while (base < a.dst - 1) {
			unsigned long long const x = s0 + s1;
			xoro_advance(s0, s1);
			double const u = __dmul_rn(__ull2double_rn((x >> 11) + 1), 0x1p-53); // (0, 1]
			double const g = __dmul_rn(spice::detail::glibc::log_with_table(u, tab), a.inv_log1mp);
			long long gap  = 1 + (g < 4.0e9 ? static_cast<long long>(g) : 4000000000ll);
			long long scan = gap; // inclusive scan over the lanes
#pragma unroll
			for (int off = 1; off < 32; off <<= 1) {
				long long const o = __shfl_up_sync(0xffffffffu, scan, off);
				if (lane >= off)
					scan += o;
			}
			long long const t  = base + scan;
			bool const keep    = t < a.dst && t >= a.col_lo && t < a.col_hi;
			unsigned const m   = __ballot_sync(0xffffffffu, keep);
			if (kWrite && keep)
				a.neighbors[out0 + kept + __popc(m & ((1u << lane) - 1))] = static_cast<int>(t - a.col_lo);
			kept += __popc(m);
			base += __shfl_sync(0xffffffffu, scan, 31);
		}
		if (!kWrite && lane == 0)
			a.degree[r] = kept;
	}
}

// ---- tiles ----
constexpr int kTileWarps = 8;    // units per tile = warps per CTA (4 measures the same)
constexpr int kUnitCap   = 1536; // entries a warp can hold: the mean is < 1024, sigma < 32
constexpr int kPipedCap  = 1280; // ... in each of its two buffers in the pipelined kernel (mean + 8 sigma: a unit that does not fit
                                 // sends the whole generation back to the one-buffer kernel)
constexpr int kTabSmem   = 4096; // table entries (pairs) kept in shared memory (p >= 0.0054); larger tables are read through L1

struct tile_args {
	long long src, dst, col_lo, col_hi;
	unsigned long long seed_lo, seed_hi;
	int block_log2;             // B = 1 << block_log2
	long long nblocks;          // blocks per row, ceil(dst / B)
	long long b_lo, nb_local;   // blocks that overlap [col_lo, col_hi)
	long long units, tiles;
	float s;                    // -1 / log2(1 - p): gaps per halving of u
	int K;                      // T[1 .. K], T[K] = 0, T[0] = 2^32 - 1 (never read as a threshold)
	uint2 const* tab;           // [K] (T[c], T[c + 1])
	unsigned long long* desc;   // [tiles] status << 62 | count
	unsigned long long* ticket;
	long long* offsets;
	int* neighbors;
	long long cap;              // entries allocated in neighbors
	int* flags;                 // 1: a unit did not fit kUnitCap, 2: neighbors too small
	int experiment;             // SPICE_GEN_EXPERIMENT (measurements only, output is wrong): 1 = no look-back (tiles placed by index)
};

// #{k >= 1 : u < T[k]} from an estimate that is wrong once in ~1e5 draws
template <bool kSmemTab>
__device__ __forceinline__ uint2 tab_at(uint2 const* tab, unsigned tab_s, int c) {
	if constexpr (kSmemTab) {
		uint2 t;
		asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(t.x), "=r"(t.y) : "r"(tab_s + 8u * static_cast<unsigned>(c)));
		return t;
	} else
		return __ldg(tab + c);
}

template <bool kSmemTab>
__device__ __noinline__ int gap_fix(unsigned u, uint2 const* tab, unsigned tab_s, int K, int c) {
	while (c + 1 < K && tab_at<kSmemTab>(tab, tab_s, c).y > u)
		c++;
	while (c > 0 && tab_at<kSmemTab>(tab, tab_s, c).x <= u)
		c--;
	return c;
}

// the estimate floor((32 - log2(u + 1)) s) and whether the table confirms it
template <bool kSmemTab>
__device__ __forceinline__ int gap_est(unsigned u, uint2 const* tab, unsigned tab_s, int K, float neg_s, float s32, bool& ok) {
	float lg;
	asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(__uint2float_rn(u) + 1.0f)); // the argument is in [1, 2^32]
	int const c   = min(__float2int_rz(fmaf(lg, neg_s, s32)), K - 1);
	uint2 const t = tab_at<kSmemTab>(tab, tab_s, c);
	ok            = ok && t.y <= u && (t.x > u || c == 0);
	return c;
}

__device__ __forceinline__ int scan_up(int v) { // inclusive warp scan: shfl.up hands back whether the source lane exists
#pragma unroll
	for (int off = 1; off < 32; off <<= 1)
		asm("{ .reg .s32 r; .reg .pred p; shfl.sync.up.b32 r|p, %0, %1, 0, 0xffffffff; @p add.s32 %0, %0, r; }" : "+r"(v) : "r"(off));
	return v;
}

// Entry: how a warp keeps a block-relative target until its place in the output is known (u16 while B <= 65536, p >= 2^-7)
struct gen_env {
	uint2 const* tab;
	unsigned tab_s;
	int K;
	float neg_s, s32;
};

// one unit (row r, block b), generated by one warp into buf: returns the number of kept targets
template <bool kSmemTab, class Entry, int kCap>
__device__ __forceinline__ int generate_unit(tile_args const& a, gen_env const& g, long long v, Entry* buf, int lane, long long& r, long long& b) {
	util::seed_seq const seed(UInt128{a.seed_lo, a.seed_hi});
	Entry* const buf4     = buf + lane * 4;
	long long const B     = 1ll << a.block_log2;
	r                     = v / a.nb_local;
	b                     = a.b_lo + (v - r * a.nb_local);
	long long const blk0  = b << a.block_log2;
	int const bsize       = static_cast<int>(min(B, a.dst - blk0));
	int const lo          = static_cast<int>(max(a.col_lo - blk0, 0ll));
	int const hi          = static_cast<int>(min(a.col_hi - blk0, static_cast<long long>(bsize)));
	UInt128 const st      = seed.stream(static_cast<UInt>(r * a.nblocks + b) * 32 + static_cast<UInt>(lane)).seed();
	unsigned long long s0 = st.lo, s1 = st.hi;
	int pos               = -1; // the last target drawn so far, relative to the block
	int n                 = 0;
	while (pos < hi - 1) {
		unsigned long long const x0 = s0 + s1;
		xoro_advance(s0, s1);
		unsigned long long const x1 = s0 + s1;
		xoro_advance(s0, s1);
		unsigned const u0 = static_cast<unsigned>(x0 >> 32), u1 = static_cast<unsigned>(x0), u2 = static_cast<unsigned>(x1 >> 32),
		               u3 = static_cast<unsigned>(x1);
		bool ok = true;
		int c0  = gap_est<kSmemTab>(u0, g.tab, g.tab_s, g.K, g.neg_s, g.s32, ok);
		int c1  = gap_est<kSmemTab>(u1, g.tab, g.tab_s, g.K, g.neg_s, g.s32, ok);
		int c2  = gap_est<kSmemTab>(u2, g.tab, g.tab_s, g.K, g.neg_s, g.s32, ok);
		int c3  = gap_est<kSmemTab>(u3, g.tab, g.tab_s, g.K, g.neg_s, g.s32, ok);
		if (!ok) [[unlikely]] {
			c0 = gap_fix<kSmemTab>(u0, g.tab, g.tab_s, g.K, c0);
			c1 = gap_fix<kSmemTab>(u1, g.tab, g.tab_s, g.K, c1);
			c2 = gap_fix<kSmemTab>(u2, g.tab, g.tab_s, g.K, c2);
			c3 = gap_fix<kSmemTab>(u3, g.tab, g.tab_s, g.K, c3);
		}
		int const l1 = c0 + 1, l2 = l1 + c1 + 1, l3 = l2 + c2 + 1, l4 = l3 + c3 + 1;
		int const incl = scan_up(l4);
		int const tot  = __shfl_sync(0xffffffffu, incl, 31);
		int const base = pos + incl - l4;
		int const t0 = base + l1, t1 = base + l2, t2 = base + l3, t3 = base + l4;
		if (n + 128 > kCap) {
			if (lane == 0)
				atomicOr(a.flags, 1);
			break;
		}
		if (pos + 1 >= lo && pos + tot < hi) { // every target of the iteration is kept
			if ((n & 3) == 0) {
				if constexpr (sizeof(Entry) == 2)
					*reinterpret_cast<uint2*>(buf4 + n) = make_uint2(static_cast<unsigned>(t0) | static_cast<unsigned>(t1) << 16,
					                                                 static_cast<unsigned>(t2) | static_cast<unsigned>(t3) << 16);
				else
					*reinterpret_cast<int4*>(buf4 + n) = make_int4(t0, t1, t2, t3);
			} else { // only behind a first iteration that dropped targets in front of this rank's columns
				buf4[n]     = static_cast<Entry>(t0);
				buf4[n + 1] = static_cast<Entry>(t1);
				buf4[n + 2] = static_cast<Entry>(t2);
				buf4[n + 3] = static_cast<Entry>(t3);
			}
			n += 128;
		} else {
			bool const k0 = t0 >= lo && t0 < hi, k1 = t1 >= lo && t1 < hi, k2 = t2 >= lo && t2 < hi, k3 = t3 >= lo && t3 < hi;
			int const mine = k0 + k1 + k2 + k3;
			int const cs   = scan_up(mine);
			int at = n + cs - mine;
			if (k0)
				buf[at++] = static_cast<Entry>(t0);
			if (k1)
				buf[at++] = static_cast<Entry>(t1);
			if (k2)
				buf[at++] = static_cast<Entry>(t2);
			if (k3)
				buf[at++] = static_cast<Entry>(t3);
			n += __shfl_sync(0xffffffffu, cs, 31);
		}
		pos += tot;
	}
	return n;
}

// a unit's targets from shared memory to their place in the output, and the row's offset if the unit starts the row
template <class Entry>
__device__ __forceinline__ void write_unit(tile_args const& a, long long base, int n, long long v, long long r, long long b, Entry const* buf, int lane) {
	if (base + n > a.cap) {
		if (lane == 0)
			atomicOr(a.flags, 2);
	} else {
		int const add    = static_cast<int>((b << a.block_log2) - a.col_lo); // block-relative target -> local column
		int* dst         = a.neighbors + base + lane;
		Entry const* src = buf + lane;
		for (int i = lane; i < n; i += 32, dst += 32, src += 32)
			*dst = static_cast<int>(*src) + add;
	}
	if (lane == 0) {
		if (b == a.b_lo)
			a.offsets[r] = base;
		if (v == a.units - 1)
			a.offsets[a.src] = base + n;
	}
}

// decoupled look-back (one warp): publish the tile's count, add up the counts of the tiles before it back to the nearest one
// whose inclusive prefix is known, publish this tile's inclusive prefix; returns the exclusive one
__device__ __forceinline__ long long look_back(tile_args const& a, long long tile, long long tot, int lane) {
	long long excl = 0;
	if (a.experiment == 1)
		return tile * 6500;
	if (tile > 0) {
		if (lane == 0)
			atomicExch(a.desc + tile, (1ull << 62) | static_cast<unsigned long long>(tot));
		long long idx = tile - 1;
		for (;;) { // 32 predecessors at a time, nearest first
			long long const j = idx - lane;
			unsigned long long d;
			do {
				d = j >= 0 ? *reinterpret_cast<unsigned long long volatile*>(a.desc + j) : (2ull << 62);
			} while (__any_sync(0xffffffffu, (d >> 62) == 0));
			unsigned const incl_mask = __ballot_sync(0xffffffffu, (d >> 62) == 2);
			int const stop           = incl_mask ? __ffs(incl_mask) - 1 : 31;
			long long part           = lane <= stop ? static_cast<long long>(d & ((1ull << 62) - 1)) : 0;
#pragma unroll
			for (int off = 16; off > 0; off >>= 1)
				part += __shfl_xor_sync(0xffffffffu, part, off);
			excl += part;
			if (incl_mask)
				break;
			idx -= 32;
		}
	}
	if (lane == 0)
		atomicExch(a.desc + tile, (2ull << 62) | static_cast<unsigned long long>(excl + tot));
	return excl;
}

template <bool kSmemTab, class Entry>
__device__ __forceinline__ gen_env load_env(tile_args const& a, unsigned char* smem_tab) {
	gen_env g;
	if constexpr (kSmemTab) {
		uint2* const t = reinterpret_cast<uint2*>(smem_tab);
		for (int i = threadIdx.x; i < a.K; i += blockDim.x)
			t[i] = a.tab[i];
		g.tab   = t;
		g.tab_s = static_cast<unsigned>(__cvta_generic_to_shared(t));
	} else {
		g.tab   = a.tab;
		g.tab_s = 0;
	}
	g.K     = a.K;
	g.neg_s = -a.s;
	g.s32   = 32.0f * a.s;
	return g;
}

template <bool kSmemTab, class Entry>
__global__ void __launch_bounds__(kTileWarps * 32) fp_fast_tiles(tile_args a) {
	extern __shared__ __align__(16) unsigned char smem[];
	__shared__ int wcount[kTileWarps];
	__shared__ long long s_tile, s_excl;
	int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	Entry* const buf = reinterpret_cast<Entry*>(smem) + warp * kUnitCap;
	gen_env const g  = load_env<kSmemTab, Entry>(a, smem + sizeof(Entry) * kTileWarps * kUnitCap);
	for (;;) {
		// A ticket is claimed only when the CTA is free to work on it (measured: claiming the next one while the current tile is
		// in its look-back hides the round trip but makes later tiles wait for a tile nobody generates yet: 2.3 -> 6.6 ms at 1e5^2)
		__syncthreads(); // the table is loaded; the previous tile is done with s_tile / s_excl / wcount
		if (threadIdx.x == 0)
			s_tile = static_cast<long long>(atomicAdd(a.ticket, 1ull));
		__syncthreads();
		long long const tile = s_tile;
		if (tile >= a.tiles)
			return;
		long long const v = tile * kTileWarps + warp;
		int n             = 0;
		long long r = 0, b = 0;
		if (v < a.units)
			n = generate_unit<kSmemTab, Entry, kUnitCap>(a, g, v, buf, lane, r, b);
		if (lane == 0)
			wcount[warp] = n;
		__syncthreads();
		if (warp == 0) {
			int const c   = lane < kTileWarps ? wcount[lane] : 0;
			long long tot = c;
#pragma unroll
			for (int off = 16; off > 0; off >>= 1)
				tot += __shfl_xor_sync(0xffffffffu, tot, off);
			long long const excl = look_back(a, tile, tot, lane);
			if (lane == 0)
				s_excl = excl;
		}
		__syncthreads();
		if (v < a.units) {
			long long base = s_excl;
			for (int w = 0; w < warp; w++)
				base += wcount[w];
			write_unit<Entry>(a, base, n, v, r, b, buf, lane);
		}
	}
}

// The same with the look-back off the generators' path (SPICE_GEN_PIPELINED=1): 8 generator warps with two buffers each and a
// ninth warp that claims the tickets and places the tiles.  The generators go on to the next tile while the last one's
// place is being found and write it out afterwards; named barriers (ids 1..6) hand tickets, counts and places back and forth.
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <bool kSmemTab, class Entry>
__global__ void __launch_bounds__((kTileWarps + 1) * 32) fp_fast_tiles_piped(tile_args a) {
	extern __shared__ __align__(16) unsigned char smem[];
	__shared__ int s_cnt[2][kTileWarps];
	__shared__ long long s_tile[2], s_base[2][kTileWarps];
	constexpr int kAll = (kTileWarps + 1) * 32;
	constexpr int kTicket = 1, kFull = 3, kBase = 5; // + buffer index
	int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	gen_env const g = load_env<kSmemTab, Entry>(a, smem + sizeof(Entry) * 2 * kTileWarps * kPipedCap);
	__syncthreads(); // the table is loaded
	if (warp == kTileWarps) { // ---- the placing warp
		if (lane == 0)
			s_tile[0] = static_cast<long long>(atomicAdd(a.ticket, 1ull));
		__syncwarp();
		bar_arrive(kTicket + 0, kAll);
		for (int i = 0;; i++) {
			int const bf = i & 1;
			bar_sync(kFull + bf, kAll); // the counts of tile i are in
			long long const tile = s_tile[bf];
			if (tile >= a.tiles)
				return;
			// The generators want their next tile now — not earlier: a ticket held by a CTA that is still busy with the tile
			// before makes every later tile's look-back wait for it.  The round trip hides behind their stores of tile i - 1.
			if (lane == 0)
				s_tile[bf ^ 1] = static_cast<long long>(atomicAdd(a.ticket, 1ull));
			__syncwarp();
			bar_arrive(kTicket + (bf ^ 1), kAll);
			int const c = lane < kTileWarps ? s_cnt[bf][lane] : 0;
			int incl    = scan_up(c);
			long long const tot  = __shfl_sync(0xffffffffu, incl, kTileWarps - 1);
			long long const excl = look_back(a, tile, tot, lane);
			if (lane < kTileWarps)
				s_base[bf][lane] = excl + incl - c;
			__syncwarp();
			bar_arrive(kBase + bf, kAll);
		}
	}
	// ---- the generators
	Entry* const bufs = reinterpret_cast<Entry*>(smem) + warp * kPipedCap;
	long long pv = -1, pr = 0, pb = 0; // the tile before: unit (-1: none), row, block
	int pn = 0;
	for (int i = 0;; i++) {
		int const bf     = i & 1;
		Entry* const buf = bufs + bf * kTileWarps * kPipedCap;
		bar_sync(kTicket + bf, kAll);
		long long const tile = s_tile[bf];
		bool const valid     = tile < a.tiles;
		long long const v    = tile * kTileWarps + warp;
		int n                = 0;
		long long r = 0, b = 0;
		if (valid && v < a.units)
			n = generate_unit<kSmemTab, Entry, kPipedCap>(a, g, v, buf, lane, r, b);
		if (lane == 0)
			s_cnt[bf][warp] = n;
		__syncwarp();
		bar_arrive(kFull + bf, kAll);
		if (i > 0) { // the tile before has had this tile's generation to find its place
			bar_sync(kBase + (bf ^ 1), kAll);
			if (pv >= 0)
				write_unit<Entry>(a, s_base[bf ^ 1][warp], pn, pv, pr, pb, bufs + (bf ^ 1) * kTileWarps * kPipedCap, lane);
		}
		if (!valid)
			return;
		pv = v < a.units ? v : -1;
		pr = r, pb = b, pn = n;
	}
}
}

namespace {
int generate_fast_tiles(cudaStream_t stream, long long src, long long dst, double p, unsigned long long seed_lo, unsigned long long seed_hi,
                        long long col_lo, long long col_hi, result* out, std::string* err) {
	scratch S;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	GEN_CUDA(S.event(&ev0));
	GEN_CUDA(S.event(&ev1));
	// gap = 1 + #{k >= 1 : u < T[k]}, T[k] = floor(2^32 (1-p)^k): P(gap > k) = T[k] / 2^32
	std::vector<unsigned> tab{0xffffffffu};
	double const lq = std::log1p(-p); // -inf for p = 1: T[1] = 0, every gap is 1
	for (long long k = 1;; k++) {
		unsigned const t = static_cast<unsigned>(std::floor(4294967296.0 * std::exp(static_cast<double>(k) * lq)));
		tab.push_back(t);
		if (t == 0)
			break;
	}
	tile_args a{};
	a.src = src, a.dst = dst, a.col_lo = col_lo, a.col_hi = col_hi, a.seed_lo = seed_lo, a.seed_hi = seed_hi;
	a.K = static_cast<int>(tab.size()) - 1;
	if (char const* e = std::getenv("SPICE_GEN_EXPERIMENT"))
		a.experiment = std::atoi(e);
	a.s = p < 1 ? static_cast<float>(-1.0 / std::log2(1.0 - p)) : 0.0f;
	a.block_log2 = 0;
	while (static_cast<double>(1ll << a.block_log2) < 512.0 / p)
		a.block_log2++;
	long long const B = 1ll << a.block_log2;
	a.nblocks  = (dst + B - 1) / B;
	a.b_lo     = col_lo / B;
	a.nb_local = (col_hi + B - 1) / B - a.b_lo;
	a.units    = src * a.nb_local;
	a.tiles    = (a.units + kTileWarps - 1) / kTileWarps;
	double const mean = static_cast<double>(src) * static_cast<double>(col_hi - col_lo) * p;
	a.cap             = static_cast<long long>(mean + 10.0 * std::sqrt(mean * (1 - p) + 1.0)) + 1024;
	std::vector<uint2> tab2(static_cast<size_t>(a.K));
	for (int c = 0; c < a.K; c++)
		tab2[static_cast<size_t>(c)] = make_uint2(tab[static_cast<size_t>(c)], tab[static_cast<size_t>(c) + 1]);
	uint2* d_tab = nullptr;
	int* d_flags = nullptr;
	GEN_CUDA(S.alloc(&d_tab, tab2.size()));
	GEN_CUDA(S.alloc(&a.desc, static_cast<size_t>(a.tiles) + 1));
	GEN_CUDA(S.alloc(&d_flags, 4));
	GEN_CUDA(cudaMalloc(&out->offsets, sizeof(long long) * static_cast<size_t>(src + 1)));
	GEN_CUDA(cudaMemcpyAsync(d_tab, tab2.data(), sizeof(uint2) * tab2.size(), cudaMemcpyHostToDevice, stream));
	a.tab     = d_tab;
	a.ticket  = a.desc + a.tiles;
	a.offsets = out->offsets;
	a.flags   = d_flags;
	bool const smem_tab = a.K <= kTabSmem;
	bool const narrow   = a.block_log2 <= 16;
	// the pipelined kernel (two buffers per warp, a ninth warp places the tiles) unless SPICE_GEN_PIPELINED=0; its buffers hold
	// mean + 8 sigma, a unit beyond that sends the generation back to the one-buffer kernel (mean + 16 sigma)
	char const* const pe = std::getenv("SPICE_GEN_PIPELINED");
	bool piped           = !(pe && pe[0] == '0') && a.experiment == 0;
	int dev = 0, sms = 148;
	GEN_CUDA(cudaGetDevice(&dev));
	GEN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
	for (int attempt = 0; attempt < 4; attempt++) {
		int const threads = (kTileWarps + (piped ? 1 : 0)) * 32;
		auto const kernel = piped ? (smem_tab ? (narrow ? fp_fast_tiles_piped<true, unsigned short> : fp_fast_tiles_piped<true, int>)
		                                      : (narrow ? fp_fast_tiles_piped<false, unsigned short> : fp_fast_tiles_piped<false, int>))
		                          : (smem_tab ? (narrow ? fp_fast_tiles<true, unsigned short> : fp_fast_tiles<true, int>)
		                                      : (narrow ? fp_fast_tiles<false, unsigned short> : fp_fast_tiles<false, int>));
		size_t const smem = (narrow ? 2 : 4) * static_cast<size_t>(kTileWarps) * (piped ? 2 * kPipedCap : kUnitCap) +
		                    (smem_tab ? sizeof(uint2) * static_cast<size_t>(a.K) : 0);
		int per_sm = 1;
		GEN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
		GEN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
		int const grid = static_cast<int>(std::min<long long>(a.tiles, static_cast<long long>(sms) * std::max(per_sm, 1)));
		if (!out->neighbors)
			GEN_CUDA(cudaMalloc(&out->neighbors, sizeof(int) * static_cast<size_t>(a.cap + 8)));
		a.neighbors = out->neighbors;
		GEN_CUDA(cudaEventRecord(ev0, stream));
		GEN_CUDA(cudaMemsetAsync(a.desc, 0, sizeof(unsigned long long) * (static_cast<size_t>(a.tiles) + 1), stream));
		GEN_CUDA(cudaMemsetAsync(d_flags, 0, sizeof(int), stream));
		kernel<<<grid, threads, smem, stream>>>(a);
		GEN_CUDA(cudaGetLastError());
		GEN_CUDA(cudaEventRecord(ev1, stream));
		long long edges = 0;
		int flags       = 0;
		GEN_CUDA(cudaMemcpyAsync(&edges, out->offsets + src, sizeof(long long), cudaMemcpyDeviceToHost, stream));
		GEN_CUDA(cudaMemcpyAsync(&flags, d_flags, sizeof(int), cudaMemcpyDeviceToHost, stream));
		GEN_CUDA(cudaStreamSynchronize(stream));
		out->edges = edges;
		out->launches += 1;
		cudaEventElapsedTime(&out->total_ms, ev0, ev1);
		out->rows_ms = out->total_ms;
		if (piped && std::getenv("SPICE_GEN_FORCE_FALLBACK")) // tests: take the path below
			flags |= 1;
		if ((flags & 1) && piped) { // a unit beyond mean + 8 sigma: once more with the larger buffers
			piped = false;
			continue;
		}
		if (flags & 1) {
			if (err)
				*err = "counter-based generator: a unit exceeded its capacity (mean + 16 sigma)";
			return 4; // SPICE_ERR_INTERNAL
		}
		if (!(flags & 2))
			return 0;
		cudaFree(out->neighbors); // mean + 10 sigma was too small: the run has counted the edges, repeat with that size
		out->neighbors = nullptr;
		a.cap          = edges;
	}
	if (err)
		*err = "counter-based generator: the output did not fit its exact size";
	return 4;
}
}

int generate_fixed_probability_fast(void* stream_, long long src, long long dst, double p, unsigned long long seed_lo, unsigned long long seed_hi,
                                    long long col_lo, long long col_hi, result* out, std::string* err) {
	auto stream = static_cast<cudaStream_t>(stream_);
	*out        = result{};
	if (src > 0 && dst > 0 && col_hi > col_lo && p >= 0x1p-13 && std::getenv("SPICE_GEN_FAST_LOG_PATH") == nullptr)
		return generate_fast_tiles(stream, src, dst, p, seed_lo, seed_hi, col_lo, col_hi, out, err);
	scratch S;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	GEN_CUDA(S.event(&ev0));
	GEN_CUDA(S.event(&ev1));
	GEN_CUDA(cudaEventRecord(ev0, stream));
	GEN_CUDA(cudaMalloc(&out->offsets, sizeof(long long) * static_cast<size_t>(src + 1)));
	GEN_CUDA(cudaMemsetAsync(out->offsets, 0, sizeof(long long) * static_cast<size_t>(src + 1), stream));
	if (src == 0 || dst == 0 || p == 0 || col_hi <= col_lo) {
		GEN_CUDA(cudaMalloc(&out->neighbors, sizeof(int) * 8));
		GEN_CUDA(cudaStreamSynchronize(stream));
		return 0;
	}
	{
		static bool uploaded[64] = {};
		static std::mutex upload_mutex;
		std::lock_guard<std::mutex> lock(upload_mutex);
		int dev = 0;
		GEN_CUDA(cudaGetDevice(&dev));
		if (dev >= 0 && dev < 64 && !uploaded[dev]) {
			GEN_CUDA(cudaMemcpyToSymbol(g_log_tab, spice::detail::glibc::log_tab, sizeof(g_log_tab)));
			uploaded[dev] = true;
		}
	}
	long long* degree = nullptr;
	void* tmp         = nullptr;
	GEN_CUDA(S.alloc(&degree, static_cast<size_t>(src + 1)));
	GEN_CUDA(cudaMemsetAsync(degree, 0, sizeof(long long) * static_cast<size_t>(src + 1), stream));
	fast_args fa{src, dst, col_lo, col_hi, seed_lo, seed_hi, p < 1 ? 1.0 / std::log1p(-p) : -0.0, degree, nullptr, nullptr};
	int const grid = static_cast<int>(std::min<long long>((src + 7) / 8, 148 * 16));
	fp_fast_rows<false><<<grid, 256, 0, stream>>>(fa);
	size_t bytes = 0;
	GEN_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, degree, out->offsets, static_cast<long long>(src + 1), stream));
	GEN_CUDA(cudaMalloc(&tmp, std::max<size_t>(bytes, 16)));
	S.dev.push_back(tmp);
	GEN_CUDA(cub::DeviceScan::ExclusiveSum(tmp, bytes, degree, out->offsets, static_cast<long long>(src + 1), stream));
	long long edges = 0;
	GEN_CUDA(cudaMemcpyAsync(&edges, out->offsets + src, sizeof(long long), cudaMemcpyDeviceToHost, stream));
	GEN_CUDA(cudaStreamSynchronize(stream));
	GEN_CUDA(cudaMalloc(&out->neighbors, sizeof(int) * static_cast<size_t>(edges + 8)));
	fa.offsets   = out->offsets;
	fa.neighbors = out->neighbors;
	fp_fast_rows<true><<<grid, 256, 0, stream>>>(fa);
	GEN_CUDA(cudaGetLastError());
	GEN_CUDA(cudaEventRecord(ev1, stream));
	GEN_CUDA(cudaStreamSynchronize(stream));
	out->edges    = edges;
	out->launches = 4;
	cudaEventElapsedTime(&out->total_ms, ev0, ev1);
	(void)err;
	return 0;
}
}
