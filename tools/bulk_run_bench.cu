// Round-2 question, as a microbenchmark: does the delivery kernel's counting loop run faster when the runs LAND IN
// SHARED MEMORY (one cp.async.bulk per run, completion on an mbarrier per stage) instead of in registers (one 16-byte
// load per lane and run, 16 runs in flight, deliver.cu)?  The register pipeline drains once per half batch (the 16
// landing loads share the warp's six scoreboards; DESIGN.md §3, "where the stall samples sit") and costs 64 registers
// per thread; bulk copies are tracked by the mbarrier, not by scoreboards, and need no landing registers.
//
// Both variants read the same synthetic stream: "runs" of `groups` consecutive 16-byte groups at random 16-byte-aligned
// positions of a buffer much larger than L2 (the access pattern of tools/gather_bench.cu), every group holding four u8
// counter addresses arranged so that the 32 lanes of one counting instruction hit 32 different banks (what pack_runs
// produces), and both count them into a warp-private 2 x 5 KB tile like unit_walker does.
//
// Variants C / D (round 2) put the same two questions to CTA-shared u32 counters bumped with red.shared.add.
// Results: profiles/microbench_r02_bulk_runs.txt.  Build and run:  nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/bulk_run_bench.cu -o
// tools/build/bulk_run_bench && tools/build/bulk_run_bench [GiB]
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int kCap      = 5120;          // counters per array (one tile)
constexpr int kTileSmem = 2 * kCap + 128; // arrays A and B + dump area, as in deliver.cu

__host__ __device__ __forceinline__ unsigned long long mix(unsigned long long x) {
	x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x;
}

// group i of the stream: element e counts the byte ((i & 31) << 2 | e) of a pseudo-random 128-byte row: the lanes of a
// warp read consecutive groups, so one counting instruction touches 32 different banks
__global__ void fill(int4* buf, unsigned long long ngroups) {
	unsigned long long const i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (i >= ngroups) return;
	int const col = static_cast<int>(i & 31) << 2;
	unsigned long long const h = mix(i);
	int const rows = 2 * kCap / 128;
	buf[i] = make_int4(static_cast<int>(h % rows) * 128 + col, static_cast<int>((h >> 16) % rows) * 128 + col + 1,
	                   static_cast<int>((h >> 32) % rows) * 128 + col + 2, static_cast<int>((h >> 48) % rows) * 128 + col + 3);
}

__device__ __forceinline__ void tally(unsigned char* cnt, int4 v) {
	unsigned char const c0 = cnt[v.x], c1 = cnt[v.y], c2 = cnt[v.z], c3 = cnt[v.w];
	cnt[v.x] = c0 + 1; cnt[v.y] = c1 + 1; cnt[v.z] = c2 + 1; cnt[v.w] = c3 + 1;
}
__device__ __forceinline__ unsigned long long run_at(unsigned long long r, unsigned long long ngroups) { return ((mix(r) >> 32) * (ngroups - 64)) >> 32; }
__device__ __forceinline__ unsigned checksum(unsigned char const* cnt, int lane) {
	unsigned s = 0;
	for (int i = lane; i < 2 * kCap / 4; i += 32) s += reinterpret_cast<unsigned const*>(cnt)[i];
	return s;
}

// ---- variant A: the register pipeline of deliver.cu (kRing = 16 runs in flight per warp) ------------------------------
__device__ __forceinline__ int4 ldg_stream(void const* p) {
	int4 v;
	asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
	return v;
}
template <int FLIGHT>
__global__ void __launch_bounds__(128) count_registers(int4 const* buf, unsigned long long ngroups, int groups, int runs_per_warp, unsigned* out) {
	extern __shared__ uint4 smem4[];
	int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned char* cnt = reinterpret_cast<unsigned char*>(smem4) + warp * kTileSmem;
	for (int i = lane; i < kTileSmem / 16; i += 32) reinterpret_cast<uint4*>(cnt)[i] = make_uint4(0, 0, 0, 0);
	__syncwarp();
	unsigned long long r = (static_cast<unsigned long long>(blockIdx.x) * 4 + warp) * 0x9e3779b97f4a7c15ull;
	int4 v[FLIGHT];
#pragma unroll
	for (int j = 0; j < FLIGHT; j++) {
		v[j] = make_int4(-1, 0, 0, 0);
		if (lane < groups) v[j] = ldg_stream(buf + run_at(r++, ngroups) + lane);
	}
	for (int i = 0; i < runs_per_warp; i += FLIGHT) {
		if ((i & 127) == 0) { // u8 counters: start over before they can wrap (a merge in the real kernel)
			__syncwarp();
			for (int k = lane; k < kTileSmem / 16; k += 32) reinterpret_cast<uint4*>(cnt)[k] = make_uint4(0, 0, 0, 0);
			__syncwarp();
		}
#pragma unroll
		for (int j = 0; j < FLIGHT; j++) {
			if (v[j].x >= 0) tally(cnt, v[j]);
			__syncwarp();
			v[j] = make_int4(-1, 0, 0, 0);
			if (lane < groups) v[j] = ldg_stream(buf + run_at(r++, ngroups) + lane);
		}
	}
	__syncwarp();
	unsigned const s = checksum(cnt, lane);
	if (s == 0x12345678u) out[0] = s;
}

// ---- variant B: runs land in shared memory --------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(void const* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
	asm volatile(
	    "{\n"
	    ".reg .pred p;\n"
	    "WAIT_%=:\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
	    "@p bra DONE_%=;\n"
	    "bra WAIT_%=;\n"
	    "DONE_%=:\n"
	    "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, void const* src, unsigned bytes, unsigned long long* bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
	             "r"(bytes), "r"(smem_u32(bar))
	             : "memory");
}

// A warp owns STAGES stages of RUNS runs (512-byte slots).  Lane l < RUNS fetches run l of a stage with one bulk copy;
// lane 0 arms the stage's mbarrier with the stage's bytes.  The warp then counts a stage while the others are in flight.
template <int RUNS, int STAGES>
__global__ void __launch_bounds__(128) count_bulk(int4 const* buf, unsigned long long ngroups, int groups, int runs_per_warp, unsigned* out) {
	extern __shared__ uint4 smem4[];
	constexpr int kStageBytes = RUNS * 512;
	constexpr int kWarpBytes  = kTileSmem + STAGES * kStageBytes;
	__shared__ unsigned long long bars[4][STAGES];
	int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned char* base  = reinterpret_cast<unsigned char*>(smem4) + warp * kWarpBytes;
	unsigned char* cnt   = base;
	unsigned char* stage = base + kTileSmem; // 16-byte aligned: kTileSmem is a multiple of 16
	for (int i = lane; i < kTileSmem / 16; i += 32) reinterpret_cast<uint4*>(cnt)[i] = make_uint4(0, 0, 0, 0);
	if (lane < STAGES) mbar_init(&bars[warp][lane], 1);
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	__syncwarp();
	unsigned long long r = (static_cast<unsigned long long>(blockIdx.x) * 4 + warp) * 0x9e3779b97f4a7c15ull;
	unsigned const bytes = static_cast<unsigned>(groups) * 16;
	auto fetch = [&](int s) {
		if (lane == 0) mbar_expect_tx(&bars[warp][s], bytes * RUNS);
		__syncwarp();
		if (lane < RUNS) bulk_g2s(stage + s * kStageBytes + lane * 512, buf + run_at(r + lane, ngroups), bytes, &bars[warp][s]);
		r += RUNS;
	};
#pragma unroll
	for (int s = 0; s < STAGES; s++) fetch(s);
	int const nstage = runs_per_warp / RUNS;
	for (int it = 0; it < nstage; it++) {
		int const s = it % STAGES;
		if ((it * RUNS & 127) == 0 && it) {
			__syncwarp();
			for (int k = lane; k < kTileSmem / 16; k += 32) reinterpret_cast<uint4*>(cnt)[k] = make_uint4(0, 0, 0, 0);
			__syncwarp();
		}
		mbar_wait(&bars[warp][s], (it / STAGES) & 1);
#pragma unroll
		for (int j = 0; j < RUNS; j++) {
			if (lane < groups) tally(cnt, *reinterpret_cast<int4 const*>(stage + s * kStageBytes + j * 512 + lane * 16));
			__syncwarp();
		}
		if (it + STAGES < nstage) fetch(s); // the stage's slots are free again: every lane has read them (__syncwarp above)
	}
	__syncwarp();
	unsigned const sum = checksum(cnt, lane);
	if (sum == 0x12345678u) out[0] = sum;
}

// ---- variants C / D: CTA-shared u32 counters bumped with red.shared.add (no read-modify-write chain, no __syncwarp
// between runs, no u8 limit); C lands the runs in registers, D in shared memory with bulk copies -------------------------
// group i of the u32 stream: element e is the BYTE address of a u32 counter whose word index is congruent to the lane
__global__ void fill_u32(int4* buf, unsigned long long ngroups) {
	unsigned long long const i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (i >= ngroups) return;
	int const lane = static_cast<int>(i & 31);
	unsigned long long const h = mix(i);
	int const rows = 2 * kCap / 32;
	buf[i] = make_int4((static_cast<int>(h % rows) * 32 + lane) * 4, (static_cast<int>((h >> 16) % rows) * 32 + lane) * 4,
	                   (static_cast<int>((h >> 32) % rows) * 32 + lane) * 4, (static_cast<int>((h >> 48) % rows) * 32 + lane) * 4);
}
__device__ __forceinline__ void reds(unsigned base, int4 v) {
	asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(base + v.x) : "memory");
	asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(base + v.y) : "memory");
	asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(base + v.z) : "memory");
	asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(base + v.w) : "memory");
}
constexpr int kCntBytes32 = 2 * kCap * 4;

template <int FLIGHT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) reds_registers(int4 const* buf, unsigned long long ngroups, int groups, int runs_per_warp, unsigned* out) {
	extern __shared__ uint4 smem4[];
	int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (int i = threadIdx.x; i < kCntBytes32 / 16; i += WARPS * 32) smem4[i] = make_uint4(0, 0, 0, 0);
	__syncthreads();
	unsigned const base = smem_u32(smem4);
	unsigned long long r = (static_cast<unsigned long long>(blockIdx.x) * WARPS + warp) * 0x9e3779b97f4a7c15ull;
	int4 v[FLIGHT];
#pragma unroll
	for (int j = 0; j < FLIGHT; j++) {
		v[j] = make_int4(-1, 0, 0, 0);
		if (lane < groups) v[j] = ldg_stream(buf + run_at(r++, ngroups) + lane);
	}
	for (int i = 0; i < runs_per_warp; i += FLIGHT) {
#pragma unroll
		for (int j = 0; j < FLIGHT; j++) {
			if (v[j].x >= 0) reds(base, v[j]);
			v[j] = make_int4(-1, 0, 0, 0);
			if (lane < groups) v[j] = ldg_stream(buf + run_at(r++, ngroups) + lane);
		}
	}
	__syncthreads();
	unsigned s = 0;
	for (int i = threadIdx.x; i < kCntBytes32 / 4; i += WARPS * 32) s += reinterpret_cast<unsigned const*>(smem4)[i];
	if (s == 0x12345678u) out[0] = s;
}

template <int RUNS, int STAGES, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) reds_bulk(int4 const* buf, unsigned long long ngroups, int groups, int runs_per_warp, unsigned* out) {
	extern __shared__ uint4 smem4[];
	constexpr int kStageBytes = RUNS * 512;
	__shared__ unsigned long long bars[WARPS][STAGES];
	int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned char* stage = reinterpret_cast<unsigned char*>(smem4) + kCntBytes32 + warp * STAGES * kStageBytes;
	for (int i = threadIdx.x; i < kCntBytes32 / 16; i += WARPS * 32) smem4[i] = make_uint4(0, 0, 0, 0);
	if (lane < STAGES) mbar_init(&bars[warp][lane], 1);
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	__syncthreads();
	unsigned const base = smem_u32(smem4);
	unsigned long long r = (static_cast<unsigned long long>(blockIdx.x) * WARPS + warp) * 0x9e3779b97f4a7c15ull;
	unsigned const bytes = static_cast<unsigned>(groups) * 16;
	auto fetch = [&](int s) {
		if (lane == 0) mbar_expect_tx(&bars[warp][s], bytes * RUNS);
		__syncwarp();
		if (lane < RUNS) bulk_g2s(stage + s * kStageBytes + lane * 512, buf + run_at(r + lane, ngroups), bytes, &bars[warp][s]);
		r += RUNS;
	};
#pragma unroll
	for (int s = 0; s < STAGES; s++) fetch(s);
	int const nstage = runs_per_warp / RUNS;
	for (int it = 0; it < nstage; it++) {
		int const s = it % STAGES;
		mbar_wait(&bars[warp][s], (it / STAGES) & 1);
		int4 v[RUNS];
#pragma unroll
		for (int j = 0; j < RUNS; j++) {
			v[j] = make_int4(-1, 0, 0, 0);
			if (lane < groups) v[j] = *reinterpret_cast<int4 const*>(stage + s * kStageBytes + j * 512 + lane * 16);
		}
#pragma unroll
		for (int j = 0; j < RUNS; j++)
			if (v[j].x >= 0) reds(base, v[j]);
		__syncwarp(); // every lane has read the stage's slots
		if (it + STAGES < nstage) fetch(s);
	}
	__syncthreads();
	unsigned s = 0;
	for (int i = threadIdx.x; i < kCntBytes32 / 4; i += WARPS * 32) s += reinterpret_cast<unsigned const*>(smem4)[i];
	if (s == 0x12345678u) out[0] = s;
}

template <class K>
float time_kernel(K launch) {
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	float best = 1e30f;
	for (int rep = 0; rep < 3; rep++) {
		cudaEventRecord(e0);
		launch();
		cudaEventRecord(e1);
		CK(cudaEventSynchronize(e1));
		float ms; cudaEventElapsedTime(&ms, e0, e1);
		best = ms < best ? ms : best;
	}
	CK(cudaGetLastError());
	return best;
}

int main(int argc, char** argv) {
	size_t const bytes = (argc > 1 ? atoll(argv[1]) : 16ll) << 30;
	unsigned long long const ngroups = bytes / 16;
	int4* buf; CK(cudaMalloc(&buf, bytes));
	fill<<<static_cast<unsigned>((ngroups + 255) / 256), 256>>>(buf, ngroups);
	CK(cudaDeviceSynchronize());
	unsigned* out; CK(cudaMalloc(&out, 8));
	int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	int const runs = 4096, groups = 25; // 400-byte runs, the benchmark's average
	printf("buffer %zu GiB, %d SMs, runs of %d groups; useful GB/s = run bytes / time\n", bytes >> 30, sms, groups);
	auto report = [&](char const* name, int ctas_per_sm, size_t smem, float ms) {
		double const gb = static_cast<double>(sms) * ctas_per_sm * 4 * runs * groups * 16 / 1e9;
		printf("%-44s %2d CTAs/SM %6zu B smem/CTA %8.3f ms %8.1f GB/s\n", name, ctas_per_sm, smem, ms, gb / (ms * 1e-3));
	};
	{
		size_t const smem = 4 * kTileSmem;
		CK(cudaFuncSetAttribute(count_registers<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
		for (int c : {4, 5})
			report("registers, 16 runs in flight", c, smem, time_kernel([&] { count_registers<16><<<sms * c, 128, smem>>>(buf, ngroups, groups, runs, out); }));
	}
	auto bulk = [&](auto kernel, char const* name, int runs_per_stage, int stages) {
		size_t const smem = 4 * (static_cast<size_t>(kTileSmem) + static_cast<size_t>(stages) * runs_per_stage * 512);
		CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
		int const fit = static_cast<int>((227 * 1024) / (smem + 1024 + 4 * stages * 8));
		for (int c = fit; c >= 1 && c >= fit - 1; c--)
			report(name, c, smem, time_kernel([&] { kernel<<<sms * c, 128, smem>>>(buf, ngroups, groups, runs, out); }));
	};
	bulk(count_bulk<8, 2>, "bulk copies, 2 stages x 8 runs", 8, 2);
	bulk(count_bulk<8, 3>, "bulk copies, 3 stages x 8 runs", 8, 3);
	bulk(count_bulk<8, 4>, "bulk copies, 4 stages x 8 runs", 8, 4);
	bulk(count_bulk<16, 2>, "bulk copies, 2 stages x 16 runs", 16, 2);
	bulk(count_bulk<16, 3>, "bulk copies, 3 stages x 16 runs", 16, 3);

	// CTA-shared u32 counters + red.shared
	fill_u32<<<static_cast<unsigned>((ngroups + 255) / 256), 256>>>(buf, ngroups);
	CK(cudaDeviceSynchronize());
	auto report_w = [&](char const* name, int warps, int ctas_per_sm, size_t smem, float ms) {
		double const gb = static_cast<double>(sms) * ctas_per_sm * warps * runs * groups * 16 / 1e9;
		printf("%-52s %2d warps x %2d CTAs/SM %6zu B smem/CTA %8.3f ms %8.1f GB/s\n", name, warps, ctas_per_sm, smem, ms, gb / (ms * 1e-3));
	};
	auto reds_reg = [&](auto kernel, char const* name, int warps) {
		size_t const smem = kCntBytes32;
		CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
		int occ = 0;
		CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, warps * 32, smem));
		for (int c = occ; c >= 1 && c >= occ - 1; c--)
			report_w(name, warps, c, smem, time_kernel([&] { kernel<<<sms * c, warps * 32, smem>>>(buf, ngroups, groups, runs, out); }));
	};
	reds_reg(reds_registers<16, 4>, "red.shared, registers, 16 in flight", 4);
	reds_reg(reds_registers<16, 8>, "red.shared, registers, 16 in flight", 8);
	reds_reg(reds_registers<8, 8>, "red.shared, registers, 8 in flight", 8);
	reds_reg(reds_registers<8, 16>, "red.shared, registers, 8 in flight", 16);
	auto reds_blk = [&](auto kernel, char const* name, int warps, int runs_per_stage, int stages) {
		size_t const smem = kCntBytes32 + static_cast<size_t>(warps) * stages * runs_per_stage * 512;
		if (smem > 227 * 1024) return;
		CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
		int occ = 0;
		CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, warps * 32, smem));
		for (int c = occ; c >= 1 && c >= occ - 1; c--)
			report_w(name, warps, c, smem, time_kernel([&] { kernel<<<sms * c, warps * 32, smem>>>(buf, ngroups, groups, runs, out); }));
	};
	reds_blk(reds_bulk<8, 2, 4>, "red.shared, bulk, 2 stages x 8 runs", 4, 8, 2);
	reds_blk(reds_bulk<8, 3, 4>, "red.shared, bulk, 3 stages x 8 runs", 4, 8, 3);
	reds_blk(reds_bulk<8, 2, 8>, "red.shared, bulk, 2 stages x 8 runs", 8, 8, 2);
	reds_blk(reds_bulk<8, 3, 8>, "red.shared, bulk, 3 stages x 8 runs", 8, 8, 3);
	reds_blk(reds_bulk<8, 4, 8>, "red.shared, bulk, 4 stages x 8 runs", 8, 8, 4);
	reds_blk(reds_bulk<16, 2, 8>, "red.shared, bulk, 2 stages x 16 runs", 8, 16, 2);
	reds_blk(reds_bulk<8, 2, 16>, "red.shared, bulk, 2 stages x 8 runs", 16, 8, 2);
	reds_blk(reds_bulk<8, 3, 16>, "red.shared, bulk, 3 stages x 8 runs", 16, 8, 3);
	reds_blk(reds_bulk<4, 4, 16>, "red.shared, bulk, 4 stages x 4 runs", 16, 4, 4);
	return 0;
}
