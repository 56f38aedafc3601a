// The Vogels-Abbott network's models (reference: samples/vogels.cpp:10-59): conductance-based
// LIF with excitatory / inhibitory conductances decaying every step, static synapses.
#pragma once

#include "spice/concepts.h"

namespace spice::models::vogels {

struct lif {
	struct neuron {
		float V     = -0.06;
		float Gex   = 0;
		float Gin   = 0;
		Int32 Twait = 0;
	};

	SPICE_HD bool update(neuron& n, float const dt, auto&) const {
		Int32 const Tref    = 50;          // steps
		float const Vrest   = -0.06;       // V
		float const Vthres  = -0.05;       // V
		float const TmemInv = 1.0f / 0.02; // 1/s
		float const Eex     = 0.0;         // V
		float const Ein     = -0.08;       // V
		float const Ibg     = 0.02;        // V

		float const TexInv = 1.0f / 0.005; // 1/s
		float const TinInv = 1.0f / 0.01;  // 1/s

		bool spiked = false;
		if (--n.Twait <= 0) {
			if (n.V > Vthres) {
				n.V     = Vrest;
				n.Twait = Tref;
				spiked  = true;
			} else
				n.V += ((Vrest - n.V) + n.Gex * (Eex - n.V) + n.Gin * (Ein - n.V) + Ibg) * (dt * TmemInv);
		}

		n.Gex -= n.Gex * (dt * TexInv);
		n.Gin -= n.Gin * (dt * TinInv);

		return spiked;
	}
};
static_assert(CheckNeuron<lif>());

struct excitatory {
	float weight;
	SPICE_HD void deliver(lif::neuron& to) const { to.Gex += weight; }
};
static_assert(CheckSynapse<excitatory>());

struct inhibitory {
	float weight;
	SPICE_HD void deliver(lif::neuron& to) const { to.Gin += weight; }
};
static_assert(CheckSynapse<inhibitory>());
}
