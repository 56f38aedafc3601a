// The plastic synapse of the Brunel+ network (reference: samples/brunel+.cpp:59-99): a stateful
// synapse {W, Zpre, Zpost} with pair-based STDP, advanced lazily (update() per step with a pre- or
// post-synaptic spike, skip() across silent stretches).  Same state, parameters and expression
// trees as the reference; exp / pow go through spice/util/math.h (parity with the host libm).
#pragma once

#include "spice/concepts.h"
#include "spice/models/brunel.h"
#include "spice/util/math.h"

namespace spice::models::brunel_plus {
using spice::models::brunel::fixed_weight;
using spice::models::brunel::lif;
using spice::models::brunel::poisson;

struct plastic {
	struct synapse {
		float W     = 1e-4;
		float Zpre  = 0;
		float Zpost = 0;
	};

	SPICE_HD void deliver(synapse const& syn, lif::neuron& to) const { to.V += syn.W; }

	SPICE_HD void update(synapse& syn, float const dt, bool const pre, bool const post) const {
		float const TstdpInv = 1.0f / 0.02f;
		float const dtInv    = 1.0f / dt;

		float const w = syn.W - pre * 0.0202f * syn.W * util::math::exp(-syn.Zpost * dtInv) +
		                post * 0.01f * (1.0f - syn.W) * util::math::exp(-syn.Zpre * dtInv);
		syn.W = w < 0.0f ? 0.0f : (0.0003f < w ? 0.0003f : w); // std::clamp(w, 0.0f, 0.0003f)

		syn.Zpre += pre;
		syn.Zpost += post;

		syn.Zpre -= syn.Zpre * dt * TstdpInv;
		syn.Zpost -= syn.Zpost * dt * TstdpInv;
	}

	SPICE_HD void skip(synapse& syn, float const dt, Int const n) const {
		float const TstdpInv = 1.0f / 0.02f;

		syn.Zpre *= util::math::pow(1 - dt * TstdpInv, n);
		syn.Zpost *= util::math::pow(1 - dt * TstdpInv, n);
	}
};
static_assert(CheckSynapse<plastic>());
}
