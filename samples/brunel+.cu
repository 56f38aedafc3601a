// Brunel+ network (reference: samples/brunel+.cpp:101-125): the Brunel network with a plastic
// (pair-based STDP) E->E connection, per-synapse state {W, Zpre, Zpost} advanced lazily; 300 steps;
// prints every step's spikes as JSON (I, E, P order) like the reference.
#include "spice/models/brunel_plus.h"
#include "spice/snn.h"

#include "spike_sink.h"

using namespace spice;
using namespace spice::models::brunel_plus;

int main() {
	int const N       = 20000;
	float const dt    = 1e-4;
	float const delay = 15e-4;

	snn brunel(dt, delay, {1337});
	auto P = brunel.add_population<poisson>(N / 2);
	auto E = brunel.add_population<lif>(N * 4 / 10);
	auto I = brunel.add_population<lif>(N / 10);

	brunel.connect<fixed_weight>(P, E, fixed_probability(0.1), delay, {2.0 / N});
	brunel.connect<fixed_weight>(P, I, fixed_probability(0.1), delay, {2.0 / N});
	brunel.connect<plastic>(E, E, fixed_probability(0.1), delay);
	brunel.connect<fixed_weight>(E, I, fixed_probability(0.1), delay, {2.0 / N});
	brunel.connect<fixed_weight>(I, E, fixed_probability(0.1), delay, {-10.0 / N});
	brunel.connect<fixed_weight>(I, I, fixed_probability(0.1), delay, {-10.0 / N});

	spike_output_stream s("Brunel+");
	for (int i = 0; i < 300; i++) {
		brunel.step();
		s << I << E << P << '\n';
		pause(0.05);
	}
	return 0;
}
