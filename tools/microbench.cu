// Micro-benchmarks that size the delivery kernel's design space on the B200 at hand:
// scattered global RED.ADD (L2 atomics) vs shared-memory atomics vs plain shared-memory RMW,
// all fed by a streamed int32 index array (the CSR row), plus the streaming read alone.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__global__ void k_read(const int* __restrict__ idx, long long n, unsigned* out) {
	unsigned acc = 0;
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
		acc += idx[i];
	if (acc == 0x12345678) out[0] = acc;
}
__global__ void k_red(const int* __restrict__ idx, long long n, unsigned* cnt) {
	long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	long long const stride = (long long)gridDim.x * blockDim.x;
	for (; i + 3 * stride < n; i += 4 * stride) {
		int a = idx[i], b = idx[i + stride], c = idx[i + 2 * stride], d = idx[i + 3 * stride];
		atomicAdd(cnt + a, 1u); atomicAdd(cnt + b, 1u); atomicAdd(cnt + c, 1u); atomicAdd(cnt + d, 1u);
	}
	for (; i < n; i += stride) atomicAdd(cnt + idx[i], 1u);
}
// each block owns a tile of `tile` counters in shared memory; indices are pre-bucketed so block b's
// slice of idx only holds targets of tile b (relative). mode 0: smem atomics, 1: plain RMW per warp-private subtile
template <int MODE>
__global__ void k_smem(const int* __restrict__ idx, long long per_block, int tile, unsigned* out) {
	extern __shared__ unsigned s[];
	for (int i = threadIdx.x; i < tile; i += blockDim.x) s[i] = 0;
	__syncthreads();
	const int* my = idx + (long long)blockIdx.x * per_block;
	if (MODE == 0) {
		for (long long i = threadIdx.x; i < per_block; i += blockDim.x) atomicAdd(&s[my[i]], 1u);
	} else {
		// warp-private sub-tile: warp w only sees indices in [w*sub, (w+1)*sub): emulate by folding
		int const warps = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
		int const sub = tile / warps;
		long long const per_warp = per_block / warps;
		const int* mine = my + w * per_warp;
		for (long long i = lane; i < per_warp; i += 32) { int t = w * sub + (mine[i] % sub); s[t] = s[t] + 1; }
	}
	__syncthreads();
	unsigned acc = 0;
	for (int i = threadIdx.x; i < tile; i += blockDim.x) acc += s[i];
	if (acc == 0x12345678) out[0] = acc;
}

int main() {
	int dev = 0; cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
	printf("device %s, %d SMs, clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
	long long const n = 1ll << 28; // 1 GiB of indices: larger than L2
	int* idx; CK(cudaMalloc(&idx, n * 4));
	unsigned* cnt; CK(cudaMalloc(&cnt, 256ll << 20));
	unsigned* out; CK(cudaMalloc(&out, 4));
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	std::vector<int> h(n);
	std::mt19937 g(1);
	auto run = [&](const char* name, auto launch, double items) {
		for (int w = 0; w < 2; w++) launch();
		CK(cudaDeviceSynchronize());
		cudaEventRecord(e0);
		int const reps = 5;
		for (int r = 0; r < reps; r++) launch();
		cudaEventRecord(e1); CK(cudaDeviceSynchronize());
		float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
		printf("%-44s %8.3f ms  %8.2f Gitems/s  %8.1f GB/s(index stream)\n", name, ms, items / ms * 1e-6, items * 4 / ms * 1e-6);
	};
	int const sms = prop.multiProcessorCount;
	// streaming read only
	for (long long i = 0; i < n; i++) h[i] = (int)(g() & 0x3ffff);
	CK(cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice));
	run("stream read int32 (1 GiB)", [&] { k_read<<<sms * 8, 512>>>(idx, n, out); }, (double)n);
	for (int bits : {18, 22, 24, 26}) { // counters: 1 MB, 16 MB, 64 MB, 256 MB
		for (long long i = 0; i < n; i++) h[i] = (int)(g() & ((1u << bits) - 1));
		CK(cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice));
		CK(cudaMemset(cnt, 0, 256ll << 20));
		char name[128]; snprintf(name, sizeof name, "global RED.ADD.u32, random over %d MB", (4 << bits) >> 20);
		run(name, [&] { k_red<<<sms * 8, 512>>>(idx, n, cnt); }, (double)n);
	}
	// ascending-with-gaps pattern like a CSR row at p = 0.02 over 2^18 targets, many rows
	{
		long long i = 0; 
		while (i < n) { int d = 0; while (i < n) { d += 1 + (int)(g() % 99); if (d >= (1 << 18)) break; h[i++] = d; } }
		CK(cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice));
		run("global RED.ADD.u32, ascending rows gap~50, 1 MB", [&] { k_red<<<sms * 8, 512>>>(idx, n, cnt); }, (double)n);
	}
	// shared memory variants: tile = 8192 counters (32 KB)
	{
		int const tile = 8192;
		for (long long i = 0; i < n; i++) h[i] = (int)(g() % tile);
		CK(cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice));
		int const blocks = sms * 4;
		long long const per_block = (n / blocks) & ~255ll;
		CK(cudaFuncSetAttribute(k_smem<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tile * 4));
		CK(cudaFuncSetAttribute(k_smem<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tile * 4));
		run("smem atomicAdd, random in 8192-tile", [&] { k_smem<0><<<blocks, 256, tile * 4>>>(idx, per_block, tile, out); }, (double)per_block * blocks);
		run("smem plain RMW, warp-private subtiles", [&] { k_smem<1><<<blocks, 256, tile * 4>>>(idx, per_block, tile, out); }, (double)per_block * blocks);
	}
	return 0;
}
