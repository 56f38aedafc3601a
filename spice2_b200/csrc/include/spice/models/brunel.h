// The Brunel network's models (reference: samples/brunel.cpp:23-72), written for this backend:
// same state, same parameters, same expression trees; SPICE_HD on the members that run inside
// the kernels, and `rng_draws` on the one model that consumes the step stream.
#pragma once

#include "spice/concepts.h"
#include "spice/util/random.h"

namespace spice::models::brunel {

// Stateless input neuron firing at 20 Hz on average: one uniform draw per neuron per step
// (samples/brunel.cpp:23-31).
struct poisson {
	static constexpr int rng_draws = 1;

	SPICE_HD bool update(float dt, auto& rng) const {
		float const firing_rate = 20; // Hz
		return util::generate_canonical<float>(rng) < (firing_rate * dt);
	}
};
static_assert(CheckNeuron<poisson>());

// Leaky integrate-and-fire with a 20-step refractory period (samples/brunel.cpp:39-62).
struct lif {
	struct neuron {
		float V   = 0;
		int Twait = 0;
	};

	SPICE_HD bool update(neuron& n, float dt, auto&) const {
		float const TmemInv = 1.0 / 0.02; // 1/s
		float const Vrest   = 0.0;        // V
		int const Tref      = 20;         // steps
		float const Vthres  = 0.02;       // V

		if (--n.Twait <= 0) {
			if (n.V > Vthres) {
				n.V     = Vrest;
				n.Twait = Tref;
				return true;
			}
			n.V += (Vrest - n.V) * (dt * TmemInv);
		}
		return false;
	}
};
static_assert(CheckNeuron<lif>());

// Static synapse: every event adds the connection's weight to the target's membrane potential
// (samples/brunel.cpp:66-70).
struct fixed_weight {
	float weight;
	SPICE_HD void deliver(lif::neuron& to) const { to.V += weight; }
};
static_assert(CheckSynapse<fixed_weight>());
}
