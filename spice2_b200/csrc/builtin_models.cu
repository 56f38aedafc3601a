// The sample models compiled into the library, so hosts without a C++/CUDA toolchain (the Python
// host mirror used by tests/ and bench.py) can build the reference's sample networks through the
// C ABI.  One translation unit = one device module: the update kernels and the synapse `apply`
// functions they call through function pointers must live together (spice/detail/model_ops.cuh).
// Compiled with -fmad=false: the functors' float expressions are evaluated as written.
#include <cstring>

#include "spice/detail/model_ops.cuh"
#include "spice/models/brunel.h"
#include "spice/models/brunel_plus.h"
#include "spice/models/vogels.h"
#include "spice_b200.h"

using namespace spice;
namespace b = spice::models::brunel;
namespace v = spice::models::vogels;

extern "C" {
spice_neuron_ops const* spice_builtin_neuron(char const* name) {
	if (!std::strcmp(name, "brunel.poisson"))
		return detail::neuron_ops<b::poisson>("brunel.poisson");
	if (!std::strcmp(name, "brunel.lif"))
		return detail::neuron_ops<b::lif>("brunel.lif");
	if (!std::strcmp(name, "vogels.lif"))
		return detail::neuron_ops<v::lif>("vogels.lif");
	return nullptr;
}

spice_synapse_ops const* spice_builtin_synapse(char const* name) {
	// the source neuron type only matters for deliver-from-to synapses; none of these is one
	if (!std::strcmp(name, "brunel.fixed_weight"))
		return detail::synapse_ops<b::fixed_weight, b::poisson, b::lif>("brunel.fixed_weight");
	if (!std::strcmp(name, "brunel+.plastic"))
		return detail::synapse_ops<models::brunel_plus::plastic, b::lif, b::lif>("brunel+.plastic");
	if (!std::strcmp(name, "vogels.excitatory"))
		return detail::synapse_ops<v::excitatory, v::lif, v::lif>("vogels.excitatory");
	if (!std::strcmp(name, "vogels.inhibitory"))
		return detail::synapse_ops<v::inhibitory, v::lif, v::lif>("vogels.inhibitory");
	return nullptr;
}
}

// ---- spice_selftest_libm -------------------------------------------------------------------------
namespace {
__global__ void libm_expf_kernel(float const* x, float* y, long long n) {
	long long const i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (i < n)
		y[i] = util::math::exp(x[i]);
}
__global__ void libm_pow_kernel(double const* x, long long const* e, double* y, long long n) {
	long long const i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (i < n)
		y[i] = util::math::pow(x[i], static_cast<Int>(e[i]));
}
}

extern "C" int spice_selftest_libm(int device, int kind, void const* x, int64_t const* n, void* y, int64_t count) {
	if (cudaSetDevice(device) != cudaSuccess)
		return SPICE_ERR_NO_DEVICE;
	size_t const w = kind == 0 ? 4 : 8;
	void *dx = nullptr, *dy = nullptr, *dn = nullptr;
	bool ok = cudaMalloc(&dx, w * count + 8) == cudaSuccess && cudaMalloc(&dy, w * count + 8) == cudaSuccess &&
	          cudaMalloc(&dn, 8 * count + 8) == cudaSuccess;
	ok = ok && cudaMemcpy(dx, x, w * count, cudaMemcpyHostToDevice) == cudaSuccess;
	if (ok && kind == 1)
		ok = cudaMemcpy(dn, n, 8 * count, cudaMemcpyHostToDevice) == cudaSuccess;
	if (ok && count > 0) {
		unsigned const blocks = static_cast<unsigned>((count + 255) / 256);
		if (kind == 0)
			libm_expf_kernel<<<blocks, 256>>>(static_cast<float const*>(dx), static_cast<float*>(dy), count);
		else
			libm_pow_kernel<<<blocks, 256>>>(static_cast<double const*>(dx), static_cast<long long const*>(dn), static_cast<double*>(dy), count);
		ok = cudaDeviceSynchronize() == cudaSuccess && cudaMemcpy(y, dy, w * count, cudaMemcpyDeviceToHost) == cudaSuccess;
	}
	cudaFree(dx);
	cudaFree(dy);
	cudaFree(dn);
	return ok ? SPICE_OK : SPICE_ERR_CUDA;
}
