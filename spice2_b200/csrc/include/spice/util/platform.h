// Host/device portability layer for the B200 backend.
//
// User models (neuron / synapse functors) are plain C++ structs, exactly as with the reference
// (spice/concepts.h), with one source-level delta: member functions that run inside the
// simulation kernels carry SPICE_HD so nvcc emits them for sm_100a as well as for the host.
#pragma once

#include <cstdint>

#if defined(__CUDACC__)
	#define SPICE_HD __host__ __device__
	#define SPICE_D __device__
	#define SPICE_FORCEINLINE __forceinline__
#else
	#define SPICE_HD
	#define SPICE_D
	#define SPICE_FORCEINLINE inline __attribute__((always_inline))
#endif

// Integer vocabulary of the reference (spice/util/stdint.h:5-12), at global scope as there.
using Int8   = std::int8_t;
using Int16  = std::int16_t;
using Int32  = std::int32_t;
using Int    = std::int64_t;
using UInt8  = std::uint8_t;
using UInt16 = std::uint16_t;
using UInt32 = std::uint32_t;
using UInt   = std::uint64_t;

SPICE_HD constexpr UInt operator"" _u64(unsigned long long int x) { return UInt(x); }

// 128-bit value as two words (reference: spice/util/stdint.h:16-24)
struct UInt128 {
	UInt lo;
	UInt hi;

	SPICE_HD constexpr UInt128 operator+(UInt const n) const {
		UInt const l = lo + n;
		return UInt128{l, hi + (l < lo ? 1u : 0u)};
	}
	SPICE_HD constexpr bool operator==(UInt128 const& o) const { return lo == o.lo && hi == o.hi; }
};

namespace spice::util::fp {
// IEEE-754 round-to-nearest primitives that the compiler may not contract or reassociate.
// Device: the _rn intrinsics are never fused by nvcc.  Host: the translation units of this
// backend are compiled with -ffp-contract=off (see spice2_b200/build.py).
#if defined(__CUDA_ARCH__)
SPICE_D SPICE_FORCEINLINE double mul(double a, double b) { return __dmul_rn(a, b); }
SPICE_D SPICE_FORCEINLINE double add(double a, double b) { return __dadd_rn(a, b); }
SPICE_D SPICE_FORCEINLINE double sub(double a, double b) { return __dsub_rn(a, b); }
SPICE_D SPICE_FORCEINLINE double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
SPICE_D SPICE_FORCEINLINE float mul(float a, float b) { return __fmul_rn(a, b); }
SPICE_D SPICE_FORCEINLINE float add(float a, float b) { return __fadd_rn(a, b); }
SPICE_D SPICE_FORCEINLINE float sub(float a, float b) { return __fsub_rn(a, b); }
SPICE_D SPICE_FORCEINLINE float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
#else
inline double mul(double a, double b) { return a * b; }
inline double add(double a, double b) { return a + b; }
inline double sub(double a, double b) { return a - b; }
inline double fma(double a, double b, double c) { return __builtin_fma(a, b, c); }
inline float mul(float a, float b) { return a * b; }
inline float add(float a, float b) { return a + b; }
inline float sub(float a, float b) { return a - b; }
inline float fma(float a, float b, float c) { return __builtin_fmaf(a, b, c); }
#endif
}
