// Bit-exact restatement of glibc 2.39's single-precision `expf` (FMA variant, `__expf_fma`).
//
// Why it is here: the reference's plastic synapse calls std::exp(float) (samples/brunel+.cpp:78-79),
// which resolves to glibc libm — a third-party dependency that is not part of the reference tree.
// This header restates glibc's published algorithm (sysdeps/ieee754/flt-32/e_expf.c, N = 32 table,
// from ARM optimized-routines) operation for operation, with the fused multiply-adds exactly where
// g++ placed them in Ubuntu GLIBC 2.39-0ubuntu8.5 (disassembly of the IFUNC target chosen on
// FMA+AVX2 hosts): kd' = fma(N/ln2, xd, SHIFT); r = fma(N/ln2, xd, -kd); z = fma(r, C0, C1);
// y = fma(r, C2, 1); y = fma(z, r*r, y); result = (float)(y * s).  Constants: glibc_expf_data.h.
//
// Pinned on the host against this host's libm (tests/test_host.py, tests/golden/libm_pins.npz).
#pragma once

#include "spice/detail/glibc_expf_data.h"
#include "spice/util/platform.h"

namespace spice::detail::glibc {

SPICE_HD SPICE_FORCEINLINE double expf_bits(std::uint64_t u) {
#if defined(__CUDA_ARCH__)
	return __longlong_as_double(static_cast<long long>(u));
#else
	return __builtin_bit_cast(double, u);
#endif
}

SPICE_HD SPICE_FORCEINLINE float expf_restated(float x) {
	namespace fp = spice::util::fp;
#if defined(__CUDA_ARCH__)
	std::uint32_t const ux = static_cast<std::uint32_t>(__float_as_int(x));
#else
	std::uint32_t const ux = __builtin_bit_cast(std::uint32_t, x);
#endif
	std::uint32_t const abstop = (ux >> 20) & 0x7ffu;
	if (abstop >= 0x42bu) { // |x| >= 88 or x is NaN (top12(88.0f) = 0x42b)
		if (ux == 0xff800000u)
			return 0.0f; // -inf
		if (abstop >= 0x7f8u)
			return x + x; // +inf, NaN
		if (x > 0x1.62e42ep6f)
			return x * 0x1p127f; // overflow -> +inf
		if (x < -0x1.9fe368p6f)
			return 0.0f; // underflow (glibc: 0x1p-95f * 0x1p-95f)
		if (x < -0x1.9d1d9ep6f)
			return 0x1p-149f; // glibc: 0x1.4p-75f * 0x1.4p-75f, rounded to the smallest subnormal
	}
	double const xd      = static_cast<double>(x);
	double const invln2n = expf_bits(expf_invln2n);
	double const shift   = expf_bits(expf_shift);
	double const kds     = fp::fma(invln2n, xd, shift);
#if defined(__CUDA_ARCH__)
	std::uint64_t const ki = static_cast<std::uint64_t>(__double_as_longlong(kds));
#else
	std::uint64_t const ki = __builtin_bit_cast(std::uint64_t, kds);
#endif
	double const kd       = fp::sub(kds, shift);
	double const r        = fp::fma(invln2n, xd, -kd);
#if defined(__CUDA_ARCH__)
	std::uint64_t const t = expf_tab_dev[ki & 31] + (ki << 47);
#else
	std::uint64_t const t = expf_tab[ki & 31] + (ki << 47);
#endif
	double const s        = expf_bits(t);
	double const z        = fp::fma(r, expf_bits(expf_C[0]), expf_bits(expf_C[1]));
	double const r2       = fp::mul(r, r);
	double y              = fp::fma(r, expf_bits(expf_C[2]), 1.0);
	y                     = fp::fma(z, r2, y);
	y                     = fp::mul(y, s);
	return static_cast<float>(y);
}
}
