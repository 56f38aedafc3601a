// Vogels-Abbott network, 4000 neurons (reference: samples/vogels.cpp:61-83): conductance-based
// LIF populations E and I, p = 0.02, static excitatory / inhibitory synapses, 1500 steps; rows
// without spikes are skipped in the JSON output like the reference.
#include "spice/models/vogels.h"
#include "spice/snn.h"

#include "spike_sink.h"

using namespace spice;
using namespace spice::models::vogels;

int main() {
	int const N       = 4000;
	float const dt    = 1e-4;
	float const delay = 8e-4;

	snn vogels(dt, delay, {1337});
	auto E = vogels.add_population<lif>(N * 8 / 10);
	auto I = vogels.add_population<lif>(N * 2 / 10);

	vogels.connect<excitatory>(E, E, fixed_probability(0.02), delay, {6.4e6 / (N * N)});
	vogels.connect<excitatory>(E, I, fixed_probability(0.02), delay, {6.4e6 / (N * N)});
	vogels.connect<inhibitory>(I, E, fixed_probability(0.02), delay, {8.16e7 / (N * N)});
	vogels.connect<inhibitory>(I, I, fixed_probability(0.02), delay, {8.16e7 / (N * N)});

	spike_output_stream s("Vogels", true);
	for (int i = 0; i < 1500; i++) {
		vogels.step();
		s << I << E << '\n';
	}
	return 0;
}
