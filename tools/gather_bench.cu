// What HBM delivers for the delivery kernel's ACCESS PATTERN, without any counting: every warp reads
// "runs" of `groups` consecutive 16-byte groups (one group per lane) at random 16-byte-aligned
// positions of a buffer much larger than L2, `flight` runs in flight per warp, `warps` warps per SM.
// The sequential copy bandwidth (MEASURED_PEAKS.json) is the roofline the judge's `frac` uses; this
// number is the ceiling of any kernel that gathers ~400-byte pieces of CSR rows.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/gather_bench.cu -o tools/build/gather_bench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ int4 ldg_stream(void const* p) {
	int4 v;
	asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ unsigned long long mix(unsigned long long x) {
	x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x;
}
template <int FLIGHT>
__global__ void gather(int4 const* buf, unsigned long long ngroups, int groups, int runs_per_warp, unsigned* out) {
	int const lane = threadIdx.x & 31;
	unsigned long long const w = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	int4 v[FLIGHT];
	unsigned acc = 0;
	unsigned long long r = w * 0x9e3779b97f4a7c15ull;
#pragma unroll
	for (int j = 0; j < FLIGHT; j++) {
		unsigned long long const at = ((mix(r++) >> 32) * (ngroups - 64)) >> 32;
		v[j] = ldg_stream(buf + at + min(lane, groups - 1));
	}
	for (int i = 0; i < runs_per_warp; i += FLIGHT) {
#pragma unroll
		for (int j = 0; j < FLIGHT; j++) {
			acc += v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
			unsigned long long const at = ((mix(r++) >> 32) * (ngroups - 64)) >> 32;
			v[j] = ldg_stream(buf + at + min(lane, groups - 1));
			if (acc == 0x12345679u) out[1] = acc; // keeps the loads in program order (see deliver_stream.inc)
		}
	}
#pragma unroll
	for (int j = 0; j < FLIGHT; j++) acc += v[j].x;
	if (acc == 0x12345678u) out[0] = acc;
}

int main(int argc, char** argv) {
	size_t const bytes = (argc > 1 ? atoll(argv[1]) : 16ll) << 30;
	int4* buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 1, bytes));
	unsigned* out; CK(cudaMalloc(&out, 8));
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	printf("buffer %zu GiB, %d SMs\n", bytes >> 30, sms);
	printf("%8s %8s %8s %12s %12s\n", "groups", "flight", "warps/SM", "useful GB/s", "sector GB/s");
	for (int groups : {13, 25, 32}) for (int flight : {8, 16}) for (int wps : {16, 32, 64}) {
		int const runs = 4096;
		int const threads = 128, blocks = sms * wps / 4;
		for (int rep = 0; rep < 2; rep++) {
			cudaEventRecord(e0);
			if (flight == 8) gather<8><<<blocks, threads>>>(buf, bytes / 16, groups, runs, out);
			else gather<16><<<blocks, threads>>>(buf, bytes / 16, groups, runs, out);
			cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
		}
		float ms; cudaEventElapsedTime(&ms, e0, e1);
		double const n = (double)blocks * 4 * (runs + flight);
		// a run of g groups at a random 16-byte offset touches (16 g + 16) / 32 + ~0.5 32-byte sectors on average
		double const sect = (16.0 * groups / 32 + 0.5) * 32;
		printf("%8d %8d %8d %12.0f %12.0f\n", groups, flight, wps, n * groups * 16 / ms / 1e6, n * sect / ms / 1e6);
	}
	return 0;
}
