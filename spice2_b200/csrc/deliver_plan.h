// Pure arithmetic of the delivery kernel (deliver.cu), shared with the host tests
// (tests/test_host.py::test_delivery_tickets_cover_every_unit, ::test_counter_address_rotation).
#pragma once

#ifdef __CUDACC__
#define SPICE_PLAN_HD __host__ __device__ __forceinline__
#else
#define SPICE_PLAN_HD inline
#endif

namespace spice::deliver {

// ---- counter addresses ------------------------------------------------------------------------------
// Every local target t of a tile (t < cap, cap a multiple of 128) has TWO u32 counters in the CTA's shared memory:
//   array A: word t                (bank t & 31)
//   array B: word cap + rotw_fwd(t) (bank (t + row + row / 32) & 31, row = t / 32: the bank field of t rotated by
//                                    a per-row amount, so targets that share a bank in A mostly do not in B)
//   dump:    words 2 cap .. 2 cap + 31 (padding entries; never read)
// A stream entry is the BYTE offset of one of them (4 x the word index), which is what red.shared takes.
SPICE_PLAN_HD int rotw_amount(int row) { return (row + (row >> 5)) & 31; }
SPICE_PLAN_HD int rotw_fwd(int t) { return (t & ~31) | ((t + rotw_amount(t >> 5)) & 31); }
SPICE_PLAN_HD int rotw_inv(int u) { return (u & ~31) | ((u - rotw_amount(u >> 5)) & 31); }

// ---- work tickets ---------------------------------------------------------------------------------------
// A unit = (connection, step of the window, tile); units are numbered connection by connection in schedule order
// (heaviest first), inside a connection step-major: unit = tile_prefix[c] * nsteps + s * tiles[c] + k.
// A CTA works through a private sequence of units: the first kStaticUnits of it are fixed by its index
// (cta, cta + grid, ...), so their spike lists can be prefetched before anything has been claimed; the rest are
// claimed from a global counter, one per finished unit, kStaticUnits units ahead of the one being counted.
constexpr int kStaticUnits = 4;  // = how far ahead of the unit being counted a warp may prefetch, plus one
constexpr int kTicketRing  = 8;  // slots of the CTA's ticket ring (> kStaticUnits)

SPICE_PLAN_HD unsigned static_ticket(unsigned cta, unsigned grid, unsigned j) { return cta + j * grid; }
SPICE_PLAN_HD unsigned dynamic_ticket(unsigned grid, unsigned claimed) { return kStaticUnits * grid + claimed; }

struct unit_pos {
	int c, s, k;
};
// tile_prefix[c] = tiles of the connections scheduled before c (tile_prefix[nconns] = total_tiles)
template <class Prefix>
SPICE_PLAN_HD unit_pos locate_unit(unsigned unit, Prefix const& tile_prefix, int nconns, int nsteps) {
	int c = 0;
	while (c + 1 < nconns && static_cast<unsigned>(tile_prefix[c + 1]) * nsteps <= unit)
		c++;
	unsigned const local = unit - static_cast<unsigned>(tile_prefix[c]) * nsteps;
	unsigned const tiles = static_cast<unsigned>(tile_prefix[c + 1] - tile_prefix[c]);
	return unit_pos{c, static_cast<int>(local / tiles), static_cast<int>(local % tiles)};
}
}
