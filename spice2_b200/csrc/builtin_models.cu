// The sample models compiled into the library, so hosts without a C++/CUDA toolchain (the Python
// host mirror used by tests/ and bench.py) can build the reference's sample networks through the
// C ABI.  One translation unit = one device module: the update kernels and the synapse `apply`
// functions they call through function pointers must live together (spice/detail/model_ops.cuh).
// Compiled with -fmad=false: the functors' float expressions are evaluated as written.
#include <cstring>

#include "spice/detail/model_ops.cuh"
#include "spice/models/brunel.h"
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
	if (!std::strcmp(name, "vogels.excitatory"))
		return detail::synapse_ops<v::excitatory, v::lif, v::lif>("vogels.excitatory");
	if (!std::strcmp(name, "vogels.inhibitory"))
		return detail::synapse_ops<v::inhibitory, v::lif, v::lif>("vogels.inhibitory");
	return nullptr;
}
}
