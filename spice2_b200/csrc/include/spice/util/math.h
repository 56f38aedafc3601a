// Math functions user models may call from functors that run inside the simulation kernels.
//
// On the host they are the host libm's (what the reference calls).  On the device they are
// restatements chosen for parity with glibc 2.39, the libm the reference links on this platform:
//   * exp(float)            bit-exact (spice/detail/glibc_expf.h);
//   * pow(double, n), n a small non-negative integer: evaluated in double-double arithmetic and
//     rounded once — the correctly rounded power, which glibc's pow returns except in rare
//     near-halfway cases (its documented error is < 0.52 ulp); other arguments: CUDA's pow.
// A model that calls std::exp / std::pow directly still compiles for the device (CUDA's libm,
// <= 2 ulp) — its results then agree with the reference only to that tolerance.
#pragma once

#include <cmath>

#include "spice/detail/glibc_expf.h"
#include "spice/util/platform.h"

namespace spice::util::math {

SPICE_HD SPICE_FORCEINLINE float exp(float x) {
#if defined(__CUDA_ARCH__)
	return spice::detail::glibc::expf_restated(x);
#else
	return std::exp(x);
#endif
}

#if defined(__CUDA_ARCH__)
namespace detail {
struct dd {
	double hi, lo;
};
__device__ __forceinline__ dd mul(dd a, dd b) {
	double const p = __dmul_rn(a.hi, b.hi);
	double e       = __fma_rn(a.hi, b.hi, -p);
	e              = __fma_rn(a.hi, b.lo, e);
	e              = __fma_rn(a.lo, b.hi, e);
	double const s = __dadd_rn(p, e);
	return dd{s, __dadd_rn(__dsub_rn(p, s), e)};
}
}
#endif

// base^n as the reference's std::pow(float, Int) computes it: both promoted to double
SPICE_HD SPICE_FORCEINLINE double pow(double base, Int n) {
#if defined(__CUDA_ARCH__)
	if (n < 0 || n > (Int(1) << 20))
		return ::pow(base, static_cast<double>(n));
	detail::dd r{1.0, 0.0}, b{base, 0.0};
	for (Int k = n; k; k >>= 1) {
		if (k & 1)
			r = detail::mul(r, b);
		b = detail::mul(b, b);
	}
	return __dadd_rn(r.hi, r.lo);
#else
	return std::pow(base, static_cast<double>(n));
#endif
}
}
