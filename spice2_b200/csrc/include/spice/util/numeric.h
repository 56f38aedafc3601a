// Compensated summation (reference: spice/util/numeric.h:6-23).  operator+= returns the
// compensated increment y, which snn::step() uses as the step's dt (spice/src/snn.cpp:8).
// The arithmetic must not be reassociated; this backend never compiles host code with
// -ffast-math, and the primitives in fp:: are not contractible.
#pragma once

#include "spice/util/platform.h"

namespace spice::util {
template <class Real>
class kahan_sum {
public:
	SPICE_HD constexpr Real operator+=(Real delta) {
		Real const y = delta - _c;
		Real const t = _sum + y;
		_c           = (t - _sum) - y;
		_sum         = t;
		return y;
	}
	SPICE_HD constexpr operator Real() const { return _sum; }
	SPICE_HD constexpr void reset() { _sum = 0; }

private:
	Real _c   = 0;
	Real _sum = 0;
};
}
