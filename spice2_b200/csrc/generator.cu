// fixed_probability adjacency generation on the GPU, bit-exact with the reference.
//
// Reference: spice::fixed_probability::generate (spice/src/topology.cpp:80-112): ONE sequential
// xoroshiro128+ stream for the whole matrix; per source row, `noise += Exp(1/p - 1)` per draw,
// dst = index + round(noise), the row ends with the first draw whose dst >= dst_count (or when
// index reaches max_degree), and that terminating draw is discarded.  Row r therefore starts at
// stream position sum_{r'<r} (deg_r' + 1): data dependent, which is what makes the loop serial.
//
// B200 design — four kernels per chunk of the stream:
//   A  fp_values   : every stream position in parallel: jump-ahead into the stream (GF(2)
//                    polynomials, spice/util/random.h), u -> y = log(u) with the bit-exact
//                    glibc restatement (spice/detail/glibc_log.h), stored as f64; plus exact
//                    fixed-point block sums of y.
//   S  fp_segscan  : exclusive scan of the per-segment sums.
//   B  fp_orbit    : the only sequential part: one warp hops from row start to row start using
//                    the prefix sums (32 candidate end positions per ballot).  A hop is decided
//                    from interval bounds on round(noise); if the bounds disagree (probability
//                    ~1e-8 per row) one lane replays that row with the exact recurrence.
//   C  fp_rows     : one thread per row replays the reference's exact float recurrence
//                    noise = fma(y, c, noise); dst = index + trunc(noise + 0.49999999999999994)
//                    over its row, writes the row, and CHECKS that it ends exactly where the
//                    orbit said the next row starts.  By induction from row 0 the adjacency is
//                    then bit-identical to the sequential algorithm; any failed check is reported
//                    (SPICE_ERR_INTERNAL), never papered over.
//
// Floating-point forms are those of the reference build (g++ 13.3 -O2 -ffast-math
// -march=haswell), taken from its disassembly (DESIGN.md lists them).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "generator.h"
#include "spice/detail/glibc_log.h"
#include "spice/util/random.h"

namespace spice::gen {
namespace {

constexpr int kSeg        = 1024;        // stream positions generated sequentially by one lane
constexpr int kBlk        = 32;          // positions per prefix-sum block (one warp ballot)
constexpr int kWarpSpan   = 32 * kSeg;   // positions covered by one warp of fp_values
constexpr int kGroupSegs  = 4;           // segments whose checkpoints one fp_checkpoints thread steps through
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
	long long guess;           // draws that can be skipped safely before looking for a row end
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

// one step of the reference's row recurrence; returns true when the row ends at this draw
__device__ __forceinline__ bool row_step(double y, params const& P, double& noise, int& index, int& dst) {
	noise = __fma_rn(y, P.c, noise);
	dst   = index + cvttsd2si32(__dadd_rn(noise, copysign(kHalfLo, noise)));
	return (static_cast<long long>(dst) >= P.dst) | (index >= P.max_degree);
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

// ---- kernel B: the orbit of row starts -------------------------------------------------------------
struct orbit_state {
	long long row;       // next row to start
	long long pos;       // its stream position (global)
	long long exact_rows; // rows decided by the exact replay
	long long rows_done_in_chunk;
};

struct orbit_args {
	params P;
	double const* y;
	long long const* blk_local;
	long long const* seg_base;
	long long base;   // global position of y[0]
	long long limit;  // rows starting at or after this global position belong to the next chunk
	long long len;    // draws held in y (chunk + overlap)
	long long* row_start; // [src + 1], global positions
	orbit_state* st;
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

__global__ void __launch_bounds__(32) fp_orbit(orbit_args a) {
	int const lane   = threadIdx.x;
	params const& P  = a.P;
	long long row    = a.st->row;
	long long s      = a.st->pos;
	long long nexact = a.st->exact_rows;
	long long const row0 = row;

	// exact fixed-point prefix at position t (chunk frame): block base + partial block
	auto prefix_at = [&](long long tl) {
		long long const b = tl / kBlk;
		long long const t = b * kBlk + lane;
		long long q       = (t < tl) ? quantize(a.y[t], P.fix) : 0;
		for (int off = 16; off; off >>= 1)
			q += __shfl_xor_sync(0xffffffffu, q, off);
		return a.seg_base[tl / kSeg] + a.blk_local[b] + q;
	};
	long long Ps = (row < P.src && s < a.limit) ? prefix_at(s - a.base) : 0;

	while (row < P.src && s < a.limit) {
		long long const sl  = s - a.base;
		long long const sb  = sl / kBlk;
		// scan forward block by block from a safe guess
		long long blk     = (sl + P.guess) / kBlk;
		bool first_block  = true;
		long long e       = -1; // local position of the terminating draw
		long long Pnext   = 0;
		bool need_exact   = false;
		{
			// Coarse step: index + round(noise) never decreases along a row, so the row cannot have ended
			// before a draw at which it provably has not.  Test the first draw of the next 32 blocks at
			// once (one round trip instead of one per block) and start the fine scan in the block before
			// the first one that may be past the end.
			long long const B = blk + lane;
			long long const t = B * kBlk;
			bool may_be_past  = true;
			if (t < a.len) {
				may_be_past        = false;
				long long const n  = t - sl;
				if (n >= 0) {
					long long const Pt = a.seg_base[t / kSeg] + a.blk_local[B] + quantize(a.y[t], P.fix); // P(t + 1)
					double const v     = __dmul_rn(__dmul_rn(__ll2double_rn(Pt - Ps), P.inv_fix), P.c);
					long long const rh = __double2ll_rd(v + P.delta + 0.5);
					may_be_past        = (n >= P.max_degree) | (n + rh >= P.dst);
				}
			}
			unsigned const m = __ballot_sync(0xffffffffu, may_be_past);
			int const j      = m ? __ffs(m) - 1 : 32;
			if (j >= 2) {
				blk += j - 1;
				first_block = false; // its first draw provably precedes the end
			}
		}
		for (;;) {
			long long const t  = blk * kBlk + lane;
			long long const in = quantize(a.y[t], P.fix);
			long long const Pt = a.seg_base[t / kSeg] + a.blk_local[blk] + warp_incl_scan(in, lane); // P(t+1)
			long long const n  = t - sl; // index of this draw within the row
			bool lo_true = false, hi_true = false;
			if (n >= 0) {
				double const v   = __dmul_rn(__dmul_rn(__ll2double_rn(Pt - Ps), P.inv_fix), P.c);
				long long const rl = __double2ll_rd(v - P.delta + 0.5);
				long long const rh = __double2ll_rd(v + P.delta + 0.5);
				bool const cap     = n >= P.max_degree;
				lo_true            = cap | (n + rl >= P.dst);
				hi_true            = cap | (n + rh >= P.dst);
			}
			unsigned const mh = __ballot_sync(0xffffffffu, hi_true);
			unsigned const ml = __ballot_sync(0xffffffffu, lo_true);
			if (mh) {
				int const fh = __ffs(mh) - 1;
				int const fl = ml ? __ffs(ml) - 1 : 32;
				long long const tf = blk * kBlk + fh;
				if (first_block && tf > sl && fh == 0) {
					// the guess overshot (or cannot be proven not to have): rescan from the row start
					blk         = sb;
					first_block = false;
					continue;
				}
				if (fh != fl) {
					need_exact = true;
				} else {
					e     = tf;
					Pnext = __shfl_sync(0xffffffffu, Pt, fh);
				}
				break;
			}
			first_block = false;
			blk++;
		}
		if (need_exact) {
			// exact replay of this row by one lane (rare)
			long long ee = 0;
			if (lane == 0) {
				double noise = 0;
				int index = 0, dst = 0;
				long long t = sl;
				while (!row_step(a.y[t], P, noise, index, dst)) {
					index++;
					t++;
				}
				ee = t;
			}
			e = __shfl_sync(0xffffffffu, ee, 0);
			nexact++;
			Pnext = prefix_at(e + 1);
		}
		Ps = Pnext;
		row++;
		s = a.base + e + 1;
		if (lane == 0)
			a.row_start[row] = s;
	}
	if (lane == 0) {
		a.st->row                = row;
		a.st->pos                = s;
		a.st->exact_rows         = nexact;
		a.st->rows_done_in_chunk = row - row0;
	}
}

// ---- kernel C: rows ------------------------------------------------------------------------------------
struct rows_args {
	params P;
	double const* y;
	long long base;
	long long const* row_start;
	long long row_lo, row_hi;  // rows of this chunk
	long long* degree;         // [src] local (kept-column) degree, written in count mode
	long long const* offsets;  // [src + 1] output offsets (write mode)
	int* neighbors;
	long long capacity;
	int* error;                // bit 0: self-check failed, bit 1: capacity exceeded
	int write;                 // 0 = count kept columns, 1 = write
};

__global__ void __launch_bounds__(128) fp_rows(rows_args a) {
	long long const r = a.row_lo + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (r >= a.row_hi)
		return;
	params const& P    = a.P;
	long long const s  = a.row_start[r] - a.base;
	long long const en = a.row_start[r + 1] - a.base - 1; // where the orbit says the terminating draw is
	long long out      = a.write ? a.offsets[r] : 0;
	long long kept     = 0;
	double noise       = 0;
	int index = 0, dst = 0;
	long long t        = s;
	double const* y    = a.y;
	// The recurrence is sequential, its inputs are not: the draws are fetched 8 at a time, one batch
	// ahead (a chunk holds only ~1,700 rows, so a dependent load per step would leave the kernel
	// waiting on memory latency).  Reads past the row's end stay inside the chunk's overlap margin.
	constexpr int kAhead = 16;
	double cur[kAhead], nxt[kAhead];
#pragma unroll
	for (int i = 0; i < kAhead; i++)
		cur[i] = y[t + i];
	for (bool done = false; !done;) {
#pragma unroll
		for (int i = 0; i < kAhead; i++)
			nxt[i] = y[t + kAhead + i];
#pragma unroll
		for (int i = 0; i < kAhead; i++) {
			if (done)
				continue;
			if (row_step(cur[i], P, noise, index, dst)) {
				done = true;
				continue;
			}
			if (dst >= P.col_lo && dst < P.col_hi) {
				if (a.write) {
					if (out < a.capacity)
						a.neighbors[out] = dst - static_cast<int>(P.col_lo);
					out++;
				}
				kept++;
			}
			index++;
			t++;
			if (t > en) // would run past the orbit's row end: the check below reports it
				done = true;
		}
#pragma unroll
		for (int i = 0; i < kAhead; i++)
			cur[i] = nxt[i];
	}
	if (t != en)
		atomicOr(a.error, 1);
	if (a.write) {
		if (out > a.capacity)
			atomicOr(a.error, 2);
	} else
		a.degree[r] = kept;
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

int generate_fixed_probability(void* stream_, long long src, long long dst, double p, unsigned long long seed_lo,
                               unsigned long long seed_hi, long long col_lo, long long col_hi, long long chunk_draws,
                               result* out, std::string* err) {
	auto stream = static_cast<cudaStream_t>(stream_);
	*out        = result{};
	cudaEvent_t ev0 = nullptr, ev1 = nullptr, evr0 = nullptr, evr1 = nullptr;
	GEN_CUDA(cudaEventCreate(&ev0));
	GEN_CUDA(cudaEventCreate(&ev1));
	GEN_CUDA(cudaEventCreate(&evr0));
	GEN_CUDA(cudaEventCreate(&evr1));
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
	// fixed-point scale: |sum of y over one row| <= (max_degree + 1) * 36.8 must stay below 2^61
	int bits = 1;
	while (std::ldexp(1.0, bits) < (static_cast<double>(P.max_degree) + 2.0) * 37.0)
		bits++;
	int const F = std::min(44, 60 - bits);
	P.fix       = std::ldexp(1.0, F);
	P.inv_fix   = std::ldexp(1.0, -F);
	double const nmax = static_cast<double>(P.max_degree) + 2.0;
	// |float-sequential noise - exact sum|: <= n * ulp(dst)/2 ; quantisation: n * 2^-(F+1) * scale ; slack x4
	P.delta = 4.0 * (nmax * (std::ldexp(static_cast<double>(dst) + scale * 37.0 + 64.0, -52) + std::ldexp(scale + 1.0, -(F + 1)))) + 1e-9;
	{
		double const mean = static_cast<double>(dst) * p;
		double const sd   = std::sqrt(mean * (1.0 - p));
		long long g       = static_cast<long long>(mean - 5.0 * sd) - 2 * kBlk;
		P.guess           = std::max<long long>(0, std::min(g, P.max_degree - 2 * kBlk));
	}

	// expected stream length and output size
	double const mean_deg = std::min(static_cast<double>(dst) * p + 1.0, static_cast<double>(P.max_degree));
	long long const ov    = round_up(P.max_degree + 2 + 2 * kBlk, kSeg);
	long long ch          = chunk_draws > 0 ? chunk_draws : (1ll << 26);
	{
		double const est = static_cast<double>(src) * (mean_deg + 1.0) * 1.02 + 65536.0;
		if (est < static_cast<double>(ch))
			ch = static_cast<long long>(est);
	}
	ch                  = round_up(std::max<long long>(ch, kWarpSpan), kWarpSpan);
	long long const len = round_up(ch + ov, kWarpSpan);
	long long const segs = len / kSeg;
	long long const groups = (segs + kGroupSegs - 1) / kGroupSegs;

	double const kept_frac = static_cast<double>(col_hi - col_lo) / static_cast<double>(dst);
	double const exp_edges = static_cast<double>(src) * static_cast<double>(dst) * p * kept_frac;
	long long capacity     = static_cast<long long>(exp_edges + 8.0 * std::sqrt(exp_edges + 1.0) + 4096.0);
	capacity               = std::min(capacity, src * std::min(P.max_degree, col_hi - col_lo));
	capacity               = std::max<long long>(capacity, 1);

	double* y            = nullptr;
	long long *blk_local = nullptr, *seg_sum = nullptr, *seg_base = nullptr, *row_start = nullptr, *degree = nullptr;
	ulonglong2* ckpt     = nullptr;
	orbit_state* st      = nullptr;
	int* error           = nullptr;
	orbit_state* st_host = nullptr;
	bool const filtered  = !(col_lo == 0 && col_hi == dst);
	GEN_CUDA(cudaMalloc(&y, sizeof(double) * static_cast<size_t>(len)));
	GEN_CUDA(cudaMalloc(&blk_local, sizeof(long long) * static_cast<size_t>(len / kBlk)));
	GEN_CUDA(cudaMalloc(&seg_sum, sizeof(long long) * static_cast<size_t>(segs)));
	GEN_CUDA(cudaMalloc(&seg_base, sizeof(long long) * static_cast<size_t>(segs)));
	GEN_CUDA(cudaMalloc(&ckpt, sizeof(ulonglong2) * static_cast<size_t>(groups * kGroupSegs)));
	GEN_CUDA(cudaMalloc(&row_start, sizeof(long long) * static_cast<size_t>(src + 1)));
	GEN_CUDA(cudaMalloc(&st, sizeof(orbit_state)));
	GEN_CUDA(cudaMalloc(&error, sizeof(int)));
	GEN_CUDA(cudaMallocHost(&st_host, sizeof(orbit_state)));
	GEN_CUDA(cudaMalloc(&out->neighbors, sizeof(int) * static_cast<size_t>(capacity + 8))); // +8: the delivery kernel reads whole 16-byte groups
	if (filtered)
		GEN_CUDA(cudaMalloc(&degree, sizeof(long long) * static_cast<size_t>(src)));
	GEN_CUDA(cudaMemsetAsync(st, 0, sizeof(orbit_state), stream));
	GEN_CUDA(cudaMemsetAsync(error, 0, sizeof(int), stream));
	GEN_CUDA(cudaMemsetAsync(row_start, 0, sizeof(long long), stream)); // row 0 starts at position 0

	{
		static bool uploaded[64] = {};
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

	auto run_values = [&](long long chunk_base) -> int {
		ca.base_poly = to_dev(util::jump::xpow(static_cast<UInt>(chunk_base)));
		fp_checkpoints<<<static_cast<int>((groups + 127) / 128), 128, 0, stream>>>(ca);
		values_args va{ckpt, y, blk_local, seg_sum, segs, P.fix};
		fp_values<<<static_cast<int>((segs / 32 + 3) / 4), 128, 0, stream>>>(va);
		launches += 2;
		return static_cast<int>(cudaGetLastError());
	};
	auto run_rows = [&](long long chunk_base, long long rlo, long long rhi, int write) -> int {
		if (rhi <= rlo)
			return 0;
		rows_args ra{P, y, chunk_base, row_start, rlo, rhi, degree, out->offsets, out->neighbors, capacity, error, write};
		cudaEventRecord(evr0, stream);
		fp_rows<<<static_cast<int>((rhi - rlo + 127) / 128), 128, 0, stream>>>(ra);
		cudaEventRecord(evr1, stream);
		launches++;
		return static_cast<int>(cudaGetLastError());
	};

	while (row < src) {
		GEN_CUDA(static_cast<cudaError_t>(run_values(base)));
		fp_segscan<<<1, 1024, 0, stream>>>(seg_sum, seg_base, segs);
		orbit_args oa{P, y, blk_local, seg_base, base, base + ch, len, row_start, st};
		fp_orbit<<<1, 32, 0, stream>>>(oa);
		launches += 2;
		GEN_CUDA(cudaGetLastError());
		GEN_CUDA(cudaMemcpyAsync(st_host, st, sizeof(orbit_state), cudaMemcpyDeviceToHost, stream));
		GEN_CUDA(cudaStreamSynchronize(stream));
		long long const rhi = st_host->row;
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
	out->edges    = edges;
	out->launches = launches;
	out->rows_ms  = rows_ms;
	cudaEventElapsedTime(&out->total_ms, ev0, ev1);

	cudaFree(y);
	cudaFree(blk_local);
	cudaFree(seg_sum);
	cudaFree(seg_base);
	cudaFree(ckpt);
	cudaFree(row_start);
	cudaFree(st);
	cudaFree(error);
	cudaFree(degree);
	cudaFreeHost(st_host);
	cudaEventDestroy(ev0);
	cudaEventDestroy(ev1);
	cudaEventDestroy(evr0);
	cudaEventDestroy(evr1);

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
}
