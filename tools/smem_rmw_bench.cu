// Shared-memory counting primitives, conflict-free by construction (lane l always hits bank l, rows differ
// per lane), 16 warps/SM, addresses cheap to form: how many SM cycles does one warp-wide access cost?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/smem_rmw_bench.cu -o tools/build/smem_rmw_bench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)
template <int MODE>
__global__ void __launch_bounds__(128) k(int iters, unsigned* out) {
	extern __shared__ unsigned char smem[];
	int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned char* cnt = smem + warp * 10240;
	for (int i = lane; i < 10240 / 4; i += 32) reinterpret_cast<unsigned*>(cnt)[i] = 0;
	__syncwarp();
	unsigned h = threadIdx.x * 2654435761u + blockIdx.x;
	unsigned const base = static_cast<unsigned>(__cvta_generic_to_shared(cnt));
	unsigned a0[4];
	for (int q = 0; q < 4; q++) {
		h = h * 1664525u + 1013904223u;
		a0[q] = base + ((h >> 16) % 64) * 128 + lane * 4; // row random per lane (< 64), bank = lane
	}
	unsigned acc = 0;
	for (int it = 0; it < iters; it++) {
		unsigned a[4];
#pragma unroll
		for (int q = 0; q < 4; q++) a[q] = a0[q] ^ ((it & 15) << 7); // another row, same bank
		if (MODE == 0 || MODE == 4) {
			unsigned c[4];
#pragma unroll
			for (int q = 0; q < 4; q++) asm volatile("ld.shared.u8 %0, [%1];" : "=r"(c[q]) : "r"(a[q] + q) : "memory");
			if (MODE == 0) {
#pragma unroll
				for (int q = 0; q < 4; q++) asm volatile("st.shared.u8 [%0], %1;" ::"r"(a[q] + q), "r"(c[q] + 1) : "memory");
			} else acc += c[0] + c[1] + c[2] + c[3];
		} else if (MODE == 1) {
#pragma unroll
			for (int q = 0; q < 4; q++) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a[q]), "r"(1u << (8 * q)) : "memory");
		} else if (MODE == 2 || MODE == 6) {
			unsigned c[4];
#pragma unroll
			for (int q = 0; q < 4; q++) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(c[q]) : "r"(a[q]) : "memory");
			if (MODE == 2) {
#pragma unroll
				for (int q = 0; q < 4; q++) asm volatile("st.shared.u32 [%0], %1;" ::"r"(a[q]), "r"(c[q] + 1) : "memory");
			} else acc += c[0] + c[1] + c[2] + c[3];
		} else if (MODE == 5) {
#pragma unroll
			for (int q = 0; q < 4; q++) asm volatile("st.shared.u8 [%0], %1;" ::"r"(a[q] + q), "r"(it) : "memory");
		} else if (MODE == 3) {
#pragma unroll
			for (int q = 0; q < 4; q++) asm volatile("st.shared.u32 [%0], %1;" ::"r"(a[q]), "r"(it) : "memory");
		} else if (MODE == 7) { // same-row stores: lanes write consecutive words of one row (the fully coalesced case)
#pragma unroll
			for (int q = 0; q < 4; q++) asm volatile("st.shared.u32 [%0], %1;" ::"r"(base + ((it + q) & 63) * 128 + lane * 4), "r"(it) : "memory");
		}
		__syncwarp();
	}
	for (int i = lane; i < 10240 / 4; i += 32) acc += reinterpret_cast<unsigned*>(cnt)[i];
	if (acc == 0x12345678u) out[0] = acc;
}
template <int MODE>
void run(char const* name, int sms, int khz, unsigned* out) {
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	CK(cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960));
	int const iters = 20000, cps = 4;
	for (int rep = 0; rep < 2; rep++) {
		cudaEventRecord(e0);
		k<MODE><<<sms * cps, 128, 40960>>>(iters, out);
		cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
	}
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	double const groups_per_sm = (double)cps * 4 * iters, cycles = ms * 1e-3 * khz * 1e3;
	printf("%-28s %6.2f SM cycles per group of 4 warp-wide accesses\n", name, cycles / groups_per_sm);
}
int main() {
	unsigned* out; CK(cudaMalloc(&out, 8));
	int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
	run<0>("4 x (ld.u8 + st.u8)", sms, khz, out);
	run<2>("4 x (ld.u32 + st.u32)", sms, khz, out);
	run<1>("4 x red.add.u32", sms, khz, out);
	run<4>("4 x ld.u8", sms, khz, out);
	run<6>("4 x ld.u32", sms, khz, out);
	run<5>("4 x st.u8 (scattered rows)", sms, khz, out);
	run<3>("4 x st.u32 (scattered rows)", sms, khz, out);
	run<7>("4 x st.u32 (one row)", sms, khz, out);
	return 0;
}
