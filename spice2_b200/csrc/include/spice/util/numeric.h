// Compensated summation (reference: spice/util/numeric.h:6-23).  operator+= returns the
// compensated increment y, which snn::step() uses as the step's dt (spice/src/snn.cpp:8).
// The arithmetic must not be reassociated; this backend never compiles host code with
// -ffast-math, and the primitives in fp:: are not contractible.
#pragma once

#include "spice/util/platform.h"

namespace spice::util {
// Real: float or double.  The running total and the rounding error lost by the last addition are kept
// separately; every increment is first corrected by that error.
template <class Real>
class kahan_sum {
	Real _lost  = 0; // (what the last addition added) - (what it should have added)
	Real _total = 0;

public:
	// adds `delta`; returns the corrected increment (the step's dt in snn::step)
	SPICE_HD constexpr Real operator+=(Real const delta) {
		Real const corrected = delta - _lost;
		Real const next      = _total + corrected;
		_lost                = (next - _total) - corrected;
		_total               = next;
		return corrected;
	}
	SPICE_HD constexpr operator Real() const { return _total; }
	// snn.cpp:9-10 restarts the total once it reaches 1; the pending correction is kept
	SPICE_HD constexpr void reset() { _total = 0; }
};
}
