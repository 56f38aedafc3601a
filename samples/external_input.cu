// A population fed from the host (what the reference's samples/external_input.cpp demonstrates): the neuron type has a
// per-population update(dt, rng, out) instead of a per-neuron one, the runtime calls it on the host once per step and
// uploads the ids it returns into the population's spike ring on the GPU (spice_add_host_population).  Here the
// "sensor" replays a fixed four-step train over nine channels, twenty steps long; stdout is the reference's JSON.
#include <array>
#include <vector>

#include "spice/snn.h"

#include "spike_sink.h"

static constexpr Int kChannels = 9;
static constexpr Int kSteps    = 20;

// the recorded train; a file reader or an event-camera driver would sit here instead
static std::array<std::vector<Int32>, 4> const kTrain = {{{4, 5, 8}, {5}, {7, 8}, {5, 7}}};

struct replay {
	std::size_t at = 0; // position in the train

	void update(float /*dt*/, auto /*rng*/, std::vector<Int32>& fired) {
		auto const& now = kTrain[at];
		fired.insert(fired.end(), now.begin(), now.end());
		at = (at + 1) % kTrain.size();
	}
};
static_assert(spice::CheckNeuron<replay>());

int main() {
	spice::snn net(1, 1, {1337});
	auto* sensor = net.add_population<replay>(kChannels);

	spike_output_stream json("external_input");
	for (Int step = 0; step < kSteps; step++) {
		net.step();
		json << sensor << '\n';
	}
	return 0;
}
