// Bit-exact restatement of glibc 2.39's double-precision `log` (FMA variant, `__log_fma`).
//
// Why it is here: the reference's adjacency generator draws its geometric skips as
// -scale * std::log(u) (spice/util/random.h:271, called from spice/src/topology.cpp:98), and
// std::log resolves to glibc libm — a third-party dependency that is not part of the reference
// tree.  The adjacency is only bit-exact if the device evaluates the very same function, so this
// header restates glibc's published algorithm (sysdeps/ieee754/dbl-64/e_log.c, N = 128 table,
// from ARM optimized-routines) operation for operation, with the fused multiply-adds exactly
// where g++ placed them in Ubuntu GLIBC 2.39-0ubuntu8.5 (disassembly of the IFUNC target chosen
// on FMA+AVX2 hosts).  Constants: glibc_log_data.h (generated from the same libm).
//
// Domain: finite x > 0, normal (the generator only calls it with u in [2^-53, 1]).  Zero,
// negative, subnormal, inf and NaN inputs are outside the contract.
//
// Pinned by tests/test_host.py::test_glibc_log_restatement_matches_golden_pins (host build vs this host's libm on 10^7 inputs, and the
// golden vectors in tests/golden/libm_pins.npz) and on the device by tests/test_gpu_generator.py.
#pragma once

#include "spice/detail/glibc_log_data.h"
#include "spice/util/platform.h"

namespace spice::detail::glibc {

SPICE_HD SPICE_FORCEINLINE double bits_to_double(std::uint64_t u) {
#if defined(__CUDA_ARCH__)
	return __longlong_as_double(static_cast<long long>(u));
#else
	return __builtin_bit_cast(double, u);
#endif
}
SPICE_HD SPICE_FORCEINLINE std::uint64_t double_to_bits(double d) {
#if defined(__CUDA_ARCH__)
	return static_cast<std::uint64_t>(__double_as_longlong(d));
#else
	return __builtin_bit_cast(std::uint64_t, d);
#endif
}

// `tab` points at 256 words laid out as log_tab (global, shared or host memory).
SPICE_HD SPICE_FORCEINLINE double log_with_table(double x, std::uint64_t const* tab) {
	namespace fp = spice::util::fp;
	std::uint64_t const ix = double_to_bits(x);

	// |x - 1| small: ix in [asuint64(1 - 2^-4), asuint64(1 + 0x1.09p-4))
	if (ix - 0x3fee000000000000ULL < 0x0003090000000000ULL) {
		if (ix == 0x3ff0000000000000ULL)
			return 0.0;
		double const B0 = bits_to_double(log_B[0]), B1 = bits_to_double(log_B[1]),
		             B2 = bits_to_double(log_B[2]), B3 = bits_to_double(log_B[3]),
		             B4 = bits_to_double(log_B[4]), B5 = bits_to_double(log_B[5]),
		             B6 = bits_to_double(log_B[6]), B7 = bits_to_double(log_B[7]),
		             B8 = bits_to_double(log_B[8]), B9 = bits_to_double(log_B[9]),
		             B10 = bits_to_double(log_B[10]);
		double const r  = fp::sub(x, 1.0);
		double const r2 = fp::mul(r, r);
		double const r3 = fp::mul(r, r2);
		double const p1 = fp::fma(r2, B3, fp::fma(r, B2, B1));
		double const p4 = fp::fma(r2, B6, fp::fma(r, B5, B4));
		double p7       = fp::fma(r2, B9, fp::fma(r, B8, B7));
		p7              = fp::fma(r3, B10, p7);
		double y        = fp::fma(p7, r3, p4);
		y               = fp::fma(y, r3, p1);
		// hi/lo split of r + B0*r^2 (w = r*2^27; rhi = r + w - w, both steps fused)
		double const t   = fp::fma(r, 0x1p27, r);
		double const rhi = fp::fma(-0x1p27, r, t);
		double const sq  = fp::mul(rhi, rhi);
		double const rlo = fp::sub(r, rhi);
		double const hi  = fp::fma(sq, B0, r);
		double lo        = fp::fma(sq, B0, fp::sub(r, hi));
		lo               = fp::fma(fp::mul(B0, rlo), fp::add(rhi, r), lo);
		y                = fp::fma(y, r3, lo);
		return fp::add(hi, y);
	}

	// x = 2^k z, z in [OFF, 2 OFF), OFF = 0x3fe6000000000000; i = top 7 mantissa bits of z - OFF
	std::uint64_t const tmp = ix - 0x3fe6000000000000ULL;
	int const i             = static_cast<int>((tmp >> 45) & 127);
	int const k             = static_cast<int>(static_cast<std::int64_t>(tmp) >> 52);
	std::uint64_t const iz  = ix - (tmp & 0xfff0000000000000ULL);
	double const invc       = bits_to_double(tab[2 * i]);
	double const logc       = bits_to_double(tab[2 * i + 1]);
	double const z          = bits_to_double(iz);
	double const kd         = static_cast<double>(k);
	double const A0 = bits_to_double(log_A[0]), A1 = bits_to_double(log_A[1]),
	             A2 = bits_to_double(log_A[2]), A3 = bits_to_double(log_A[3]),
	             A4 = bits_to_double(log_A[4]);

	double const r  = fp::fma(z, invc, -1.0);
	double const w  = fp::fma(kd, bits_to_double(log_ln2hi), logc);
	double const hi = fp::add(r, w);
	double const r2 = fp::mul(r, r);
	double lo       = fp::add(fp::sub(w, hi), r);
	lo              = fp::fma(kd, bits_to_double(log_ln2lo), lo);
	double const r3 = fp::mul(r, r2);
	double const q  = fp::fma(fp::fma(r, A4, A3), r2, fp::fma(r, A2, A1));
	double const s  = fp::fma(r2, A0, lo);
	return fp::add(fp::fma(r3, q, s), hi);
}

#if !defined(__CUDA_ARCH__)
inline double log(double x) { return log_with_table(x, log_tab); }
#endif
}
