// Host-side GF(2) polynomial arithmetic for xoroshiro128+ jump-ahead (see spice/util/random.h).
#include "spice/util/random.h"

#include <array>
#include <mutex>

namespace spice::util::jump {
namespace {
// 256-bit GF(2) polynomial helpers for Berlekamp-Massey
using bits256 = std::array<UInt, 4>;

inline bool get(bits256 const& b, int i) { return (b[i >> 6] >> (i & 63)) & 1; }
inline void flip(bits256& b, int i) { b[i >> 6] ^= UInt(1) << (i & 63); }
inline bits256 shifted(bits256 const& b, int s) { // b * x^s
	bits256 r{};
	for (int i = 0; i + s < 256; i++)
		if (get(b, i))
			flip(r, i + s);
	return r;
}

poly compute_charpoly() {
	// 256 bits of the lowest state bit, from an arbitrary non-zero state
	xoroshiro64_128p g(0x9E3779B97F4A7C15_u64, 0xD1B54A32D192ED03_u64);
	std::array<bool, 256> a{};
	for (auto& x : a) {
		x = g.s0 & 1;
		g.advance();
	}
	// Berlekamp-Massey: connection polynomial C(x) = 1 + c1 x + ... + cL x^L
	bits256 C{}, B{};
	C[0] = B[0] = 1;
	int L = 0, m = 1;
	for (int n = 0; n < 256; n++) {
		bool d = a[n];
		for (int i = 1; i <= L; i++)
			d ^= get(C, i) && a[n - i];
		if (!d) {
			m++;
		} else if (2 * L <= n) {
			bits256 const T = C;
			bits256 const s = shifted(B, m);
			for (int w = 0; w < 4; w++)
				C[w] ^= s[w];
			L = n + 1 - L;
			B = T;
			m = 1;
		} else {
			bits256 const s = shifted(B, m);
			for (int w = 0; w < 4; w++)
				C[w] ^= s[w];
			m++;
		}
	}
	SPICE_INV(L == 128 && "xoroshiro128+ must have a degree-128 minimal polynomial");
	// P(x) = x^128 + c1 x^127 + ... + c128: coefficient of x^(128-i) is c_i
	poly p;
	for (int i = 1; i <= 128; i++)
		if (get(C, i)) {
			int const e = 128 - i;
			(e < 64 ? p.lo : p.hi) |= UInt(1) << (e & 63);
		}
	return p;
}
}

poly charpoly() {
	static poly const p = compute_charpoly();
	return p;
}

poly mulmod(poly a, poly b) {
	poly const P = charpoly();
	poly acc;
	for (int i = 127; i >= 0; i--) {
		// acc *= x (mod P)
		bool const carry = acc.hi >> 63;
		acc.hi           = (acc.hi << 1) | (acc.lo >> 63);
		acc.lo <<= 1;
		if (carry) {
			acc.lo ^= P.lo;
			acc.hi ^= P.hi;
		}
		if (b.bit(i)) {
			acc.lo ^= a.lo;
			acc.hi ^= a.hi;
		}
	}
	return acc;
}

poly powmod(poly base, UInt n) {
	poly r;
	r.lo = 1;
	while (n) {
		if (n & 1)
			r = mulmod(r, base);
		base = mulmod(base, base);
		n >>= 1;
	}
	return r;
}

poly xpow(UInt k) {
	poly x;
	x.lo = 2;
	return powmod(x, k);
}
}
