// Two populations that wake each other up in turn (what the reference's samples/ping_pong.cpp demonstrates): 1:1
// adj_list connections in both directions, a per-neuron init hook that arms one side, a one-byte neuron state and a
// stateless synapse whose deliver() is idempotent.  Ten steps; stdout is the reference's JSON.
#include "spice/snn.h"

#include "spike_sink.h"

static constexpr Int kPairs = 5'000; // 10,000 neurons
static constexpr Int kSteps = 10;

// fires once whenever it has been armed
struct relay {
	bool starts_armed;

	struct neuron {
		bool armed;
	};

	SPICE_HD void init(neuron& n, Int /*id*/, auto& /*rng*/) const { n.armed = starts_armed; }

	SPICE_HD bool update(neuron& n, float /*dt*/, auto& /*rng*/) const {
		bool const fires = n.armed;
		n.armed          = false;
		return fires;
	}
};
static_assert(spice::CheckNeuron<relay>());

struct arm {
	SPICE_HD void deliver(relay::neuron& target) const { target.armed = true; }
};
static_assert(spice::CheckSynapse<arm>());

int main() {
	spice::snn net(1, 1, {1337});
	auto* ping = net.add_population<relay>(kPairs, {true});
	auto* pong = net.add_population<relay>(kPairs, {false});

	spice::adj_list one_to_one;
	for (Int i = 0; i < kPairs; i++)
		one_to_one.connect(i, i);
	net.connect<arm>(ping, pong, one_to_one, 1);
	net.connect<arm>(pong, ping, one_to_one, 1);

	spike_output_stream json("ping-pong");
	for (Int step = 0; step < kSteps; step++) {
		net.step();
		json << ping << pong << '\n';
	}
	return 0;
}
