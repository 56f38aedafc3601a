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

#include <cmath>
#include <cstddef>
#include <cstring>
#include <initializer_list>
#include <limits>
#include <random>
#include <type_traits>

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
	// from the standard library's seed sequence: its first four 32-bit outputs are the seed (reference: random.h:145-148)
	seed_seq(std::seed_seq seq) {
		UInt32 words[4];
		seq.generate(words, words + 4);
		_seed = {static_cast<UInt>(words[0]) | static_cast<UInt>(words[1]) << 32, static_cast<UInt>(words[2]) | static_cast<UInt>(words[3]) << 32};
	}
	// SeedSequence::generate: the seed's four 32-bit words, repeated (reference: random.h:154-159)
	template <class OutputIt>
	constexpr void generate(OutputIt first, OutputIt last) const {
		UInt32 const words[4] = {static_cast<UInt32>(_seed.lo), static_cast<UInt32>(_seed.lo >> 32), static_cast<UInt32>(_seed.hi),
		                         static_cast<UInt32>(_seed.hi >> 32)};
		for (int i = 0; first != last; ++first, i = (i + 1) & 3)
			*first = words[i];
	}

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

// xoshiro128+ on four 32-bit words (reference: random.h:204-220, "xoroshiro32_128p"): 32-bit output, state = the seed's
// words in order.  Not used by the simulation loop (its one stream is the 64-bit engine above); kept for user models.
struct xoroshiro32_128p {
	using result_type = UInt32;

	UInt32 s[4] = {0, 0, 0, 0};

	SPICE_HD constexpr explicit xoroshiro32_128p(seed_seq const& seq) :
	s{static_cast<UInt32>(seq.seed().lo), static_cast<UInt32>(seq.seed().lo >> 32), static_cast<UInt32>(seq.seed().hi),
	  static_cast<UInt32>(seq.seed().hi >> 32)} {}

	SPICE_HD static constexpr UInt32 min() { return 0; }
	SPICE_HD static constexpr UInt32 max() { return ~UInt32(0); }

	SPICE_HD constexpr UInt32 operator()() {
		UInt32 const out = s[0] + s[3];
		UInt32 const t   = s[1] << 9;
		s[2] ^= s[0];
		s[3] ^= s[1];
		s[1] ^= s[2];
		s[0] ^= s[3];
		s[2] ^= t;
		s[3] = (s[3] << 11) | (s[3] >> 21);
		return out;
	}
};

// Uniform in [0,1) (LeftOpen: (0,1]) from the top mantissa-many bits of one draw; an engine narrower than Real (the 32-bit
// engine feeding a double) contributes two draws, first draw in the high half (reference: random.h:236-247).
template <class Real, bool LeftOpen = false, class Rng>
SPICE_HD constexpr Real generate_canonical(Rng& rng) {
	constexpr int digits   = std::numeric_limits<Real>::digits;
	constexpr int rng_bits = 8 * static_cast<int>(sizeof(decltype(rng())));
	constexpr int width    = rng_bits < 8 * static_cast<int>(sizeof(Real)) ? 8 * static_cast<int>(sizeof(Real)) : rng_bits;
	UInt draw              = rng();
	if constexpr (rng_bits < 8 * static_cast<int>(sizeof(Real)))
		draw = (draw << rng_bits) | rng();
	UInt const top = (draw >> (width - digits)) + (LeftOpen ? 1u : 0u);
	if constexpr (digits < 32) // the same value through a 32-bit conversion (I2F.U32 instead of the slower I2F.U64 on the device)
		return static_cast<Real>(static_cast<UInt32>(top)) / static_cast<Real>(1_u64 << digits);
	else
		return static_cast<Real>(top) / static_cast<Real>(1_u64 << digits);
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

// The reference's remaining distributions (random.h:264-330), for user models.  On the device log / sqrt / cos / sin are
// CUDA's (within an ulp of the host's libm, not bit-identical to it); the simulation loop's own bit-exact paths do not use
// these classes (the generator restates glibc's log, spice/detail/glibc_log.h).  A neuron whose update() draws through them
// declares the draws per call in rng_draws: 1 for exponential; normal draws 2 on every second call, so a model keeps the
// object per call (2 draws) or per neuron.
template <class Real>
class exponential_distribution {
public:
	SPICE_HD explicit exponential_distribution(Real scale = 1) : _scale(scale) { SPICE_PRE_HOST(scale >= 0); }
	template <class Rng>
	SPICE_HD Real operator()(Rng& rng) const {
		return -_scale * std::log(generate_canonical<Real, true>(rng)); // (0, 1] -> [0, inf)
	}

private:
	Real _scale;
};

// Box-Muller, both variates kept: a call either draws two uniforms or returns the second variate of the call before
template <class Real>
class normal_distribution {
public:
	SPICE_HD explicit normal_distribution(Real mu = 0, Real sigma = 1) : _mu(mu), _sigma(sigma) { SPICE_PRE_HOST(sigma >= 0); }
	template <class Rng>
	SPICE_HD Real operator()(Rng& rng) {
		_have = !_have;
		if (!_have)
			return _next;
		Real const radius = std::sqrt(Real(-2) * std::log(generate_canonical<Real, true>(rng)));
		// 2 pi u in double whatever Real is, as the reference's `2 * std::numbers::pi * u` evaluates
		Real const angle = static_cast<Real>(2 * 3.141592653589793238462643383279502884 * generate_canonical<Real, false>(rng));
		_next             = fp::fma(radius * std::sin(angle), _sigma, _mu);
		return fp::fma(radius * std::cos(angle), _sigma, _mu);
	}

private:
	bool _have = false; // _next holds the sine variate of the last pair
	Real _next = 0;
	Real _mu, _sigma;
};

// the normal approximation N(Np, Np(1-p)) rounded and clamped to [0, N]
template <class Integer>
class binomial_distribution {
public:
	using Real = std::conditional_t<sizeof(Integer) == 8, double, float>;
	SPICE_HD explicit binomial_distribution(Integer N, Real p) : _n(N), _normal(N * p, std::sqrt(N * p * (1 - p))) {
		SPICE_PRE_HOST(N >= 0);
		SPICE_PRE_HOST(0 <= p && p <= 1);
	}
	template <class Rng>
	SPICE_HD Integer operator()(Rng& rng) {
		Real const x       = std::round(_normal(rng));
		Integer const k    = x > Real(0) ? static_cast<Integer>(x) : Integer(0);
		return k < _n ? k : _n;
	}

private:
	Integer _n;
	normal_distribution<Real> _normal;
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
