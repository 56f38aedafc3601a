// Seeding and random number generation, host + device.
//
// Bit-compatible with the reference's spice/util/random.h: the 128-bit murmur3 finaliser chain
// that derives seeds (random.h:32-139), seed_seq (random.h:143-175), xoroshiro128+ with
// constants (24,16,37) (random.h:222-234) and the canonical / exponential maps
// (random.h:236-276).  On top of that this backend needs what a sequential CPU generator never
// did: random access into one xoroshiro stream.  xoroshiro128+'s state update is linear over
// GF(2), so the state k steps ahead is poly_k(T) s0 with poly_k = x^k mod charpoly(T); the jump
// helpers at the bottom compute such polynomials on the host and apply them on host or device.
#pragma once

#include <cstddef>
#include <cstring>
#include <initializer_list>
#include <limits>

#include "spice/util/assert.h"
#include "spice/util/platform.h"

namespace spice::util {
namespace detail {
SPICE_HD constexpr UInt rotl(UInt x, int k) { return (x << k) | (x >> (64 - k)); }

SPICE_HD constexpr UInt avalanche(UInt k) {
	k = (k ^ (k >> 33)) * 0xff51afd7ed558ccd_u64;
	k = (k ^ (k >> 33)) * 0xc4ceb9fe1a85ec53_u64;
	return k ^ (k >> 33);
}

// murmur3 x64/128 body for one 16-byte block, and its tail/finish, with the reference's
// non-zero initial hash.
struct murmur_state {
	UInt lo = 0x2E4016967F18E81_u64;
	UInt hi = 0x447567949F9AA86_u64;

	static constexpr UInt c1 = 0x87c37b91114253d5_u64;
	static constexpr UInt c2 = 0x4cf5ad432745937f_u64;

	SPICE_HD constexpr void mix_lo(UInt k) { lo ^= rotl(k * c1, 31) * c2; }
	SPICE_HD constexpr void mix_hi(UInt k) { hi ^= rotl(k * c2, 33) * c1; }
	SPICE_HD constexpr void block(UInt k1, UInt k2) {
		mix_lo(k1);
		lo = (rotl(lo, 27) + hi) * 5 + 0x52dce729;
		mix_hi(k2);
		hi = (rotl(hi, 31) + lo) * 5 + 0x38495ab5;
	}
	SPICE_HD constexpr UInt128 finish(UInt len) {
		lo ^= len;
		hi ^= len;
		lo += hi;
		hi += lo;
		lo = avalanche(lo);
		hi = avalanche(hi);
		lo += hi;
		hi += lo;
		return {lo, hi};
	}
};

inline UInt128 murmur3(void const* ptr, UInt len) {
	auto const* bytes = static_cast<unsigned char const*>(ptr);
	murmur_state m;
	UInt const nblocks = len / 16;
	for (UInt b = 0; b < nblocks; b++) {
		UInt k[2];
		std::memcpy(k, bytes + 16 * b, 16);
		m.block(k[0], k[1]);
	}
	// tail: little-endian partial words
	unsigned char tail[16] = {};
	UInt const rem         = len & 15;
	std::memcpy(tail, bytes + 16 * nblocks, rem);
	UInt k[2];
	std::memcpy(k, tail, 16);
	if (rem > 8)
		m.mix_hi(k[1]);
	if (rem > 0)
		m.mix_lo(k[0]);
	return m.finish(len);
}

SPICE_HD constexpr UInt128 murmur3(UInt128 k) {
	murmur_state m;
	m.block(k.lo, k.hi);
	return m.finish(16);
}
}

// Copy-able, fixed-size seed sequence (reference: random.h:143-175).  `seed++` hands out the
// current seed and advances the sequence by one murmur3 application.
class seed_seq {
public:
	seed_seq(std::initializer_list<UInt32> il) : _seed(detail::murmur3(il.begin(), 4 * il.size())) {
		SPICE_PRE(il.size() > 0 && "Please provide at least 1 seed to seed_seq");
	}
	seed_seq(UInt32 const* words, std::size_t n) : _seed(detail::murmur3(words, 4 * n)) {
		SPICE_PRE(n > 0 && "Please provide at least 1 seed to seed_seq");
	}
	SPICE_HD constexpr explicit seed_seq(UInt128 raw) : _seed(raw) {}

	SPICE_HD constexpr UInt128 seed() const { return _seed; }

	SPICE_HD constexpr seed_seq operator++(int) {
		seed_seq const before = *this;
		_seed                 = detail::murmur3(_seed);
		return before;
	}

	SPICE_HD constexpr seed_seq stream(UInt id) const { return seed_seq(detail::murmur3(_seed + (id + 1))); }

private:
	UInt128 _seed;
};

// xoroshiro128+ (64-bit output), state = (seed.lo, seed.hi) (reference: random.h:178-234).
struct xoroshiro64_128p {
	using result_type = UInt;

	UInt s0 = 0;
	UInt s1 = 0;

	xoroshiro64_128p() = default;
	SPICE_HD constexpr explicit xoroshiro64_128p(seed_seq const& seq) : s0(seq.seed().lo), s1(seq.seed().hi) {}
	SPICE_HD constexpr xoroshiro64_128p(UInt a, UInt b) : s0(a), s1(b) {}

	SPICE_HD static constexpr UInt min() { return 0; }
	SPICE_HD static constexpr UInt max() { return ~UInt(0); }

	SPICE_HD constexpr UInt operator()() {
		UInt const out = s0 + s1;
		advance();
		return out;
	}
	SPICE_HD constexpr void advance() {
		UInt const t = s0 ^ s1;
		s0           = detail::rotl(s0, 24) ^ t ^ (t << 16);
		s1           = detail::rotl(t, 37);
	}
};

// Uniform in [0,1) (LeftOpen: (0,1]) from the top mantissa-many bits of one 64-bit draw
// (reference: random.h:236-247; only 64-bit engines exist in this backend).
template <class Real, bool LeftOpen = false, class Rng>
SPICE_HD constexpr Real generate_canonical(Rng& rng) {
	constexpr int digits = std::numeric_limits<Real>::digits;
	UInt const draw      = rng();
	return static_cast<Real>((draw >> (64 - digits)) + (LeftOpen ? 1u : 0u)) /
	       static_cast<Real>(1_u64 << digits);
}

template <class Real, bool LeftOpen = false>
class uniform_real_distribution {
public:
	SPICE_HD constexpr explicit uniform_real_distribution(Real a = 0, Real b = 1) : _a(a), _w(b - a) {}
	template <class Rng>
	SPICE_HD Real operator()(Rng& rng) const {
		return fp::fma(generate_canonical<Real, LeftOpen>(rng), _w, _a);
	}

private:
	Real _a, _w;
};

// ---------------------------------------------------------------------------------------------
// Jump-ahead for xoroshiro128+ (new in this backend).
//
// A polynomial over GF(2) of degree < 128 is kept as a UInt128 (bit i of lo = coefficient of
// x^i, bit i of hi = coefficient of x^(64+i)).  charpoly() is the characteristic polynomial of
// the state transition, recovered once from the generator itself with Berlekamp-Massey, so the
// constants (24,16,37) live in exactly one place (xoroshiro64_128p::advance).
// ---------------------------------------------------------------------------------------------
namespace jump {
struct poly {
	UInt lo = 0, hi = 0;
	SPICE_HD constexpr bool bit(int i) const { return ((i < 64 ? lo >> i : hi >> (i - 64)) & 1) != 0; }
};

// low 128 coefficients of the (monic, degree-128) characteristic polynomial
poly charpoly();

// (a * b) mod charpoly
poly mulmod(poly a, poly b);
// x^k mod charpoly
poly xpow(UInt k);
// x^(k * 2^e)... convenience: (x^k)^n mod charpoly via square-and-multiply on the exponent n
poly powmod(poly base, UInt n);

// state after k steps from s: sum over set coefficients c_i of T^i s  (128 engine steps)
SPICE_HD inline xoroshiro64_128p apply(poly c, xoroshiro64_128p s) {
	UInt a = 0, b = 0;
	for (int i = 0; i < 128; i++) {
		UInt const m = c.bit(i) ? ~UInt(0) : 0;
		a ^= s.s0 & m;
		b ^= s.s1 & m;
		s.advance();
	}
	return {a, b};
}
}
}
