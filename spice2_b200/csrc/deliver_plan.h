// Pure arithmetic of the delivery kernels (deliver.cu), shared with the host tests
// (tests/test_host.py::test_delivery_plan_covers_every_batch, ::test_counter_address_rotation).
//
// Work items of a split delivery launch (deliver_tiles<true>):
// A unit = (connection, step, tile).  Its step holds `total` spikes = ceil(total / 32) batches; a CTA may count at most
// `per_round` batches between two merges (u8 counters), so the unit takes rounds_of() rounds.  A split launch hands
// out single rounds: (connection, step) owns tiles * rounds consecutive tickets, a unit's rounds next to each other,
// round 0 first (it stores the counters, the later rounds wait for it and add), batches spread evenly.
#pragma once

#ifdef __CUDACC__
#define SPICE_PLAN_HD __host__ __device__ __forceinline__
#else
#define SPICE_PLAN_HD inline
#endif

namespace spice::deliver {
SPICE_PLAN_HD unsigned rounds_of(bool arranged, unsigned total, unsigned per_round) {
	unsigned const nbatch = (total + 31) / 32;
	unsigned const rounds = (nbatch + per_round - 1) / per_round;
	return arranged && rounds > 1 ? rounds : 1u; // plain (multapse) connections are walked whole
}

// The two u8 counters of local target t of a tile (stream format, deliver.h): array A at byte t (shared-memory bank
// (t >> 2) & 31), array B at byte cap + rot_fwd(t), where the bank field of t is rotated by the number of t's 128-byte
// row.  Targets that share a bank in A (same bank field, different rows) therefore sit in different banks in B as
// long as the tile has at most 32 rows — what makes pack_runs' 2-choice bank balancing effective.
SPICE_PLAN_HD int rot_fwd(int t) { return (t & ~0x7c) | ((t + ((t >> 7) << 2)) & 0x7c); }
SPICE_PLAN_HD int rot_inv(int u) { return (u & ~0x7c) | ((u - ((u >> 7) << 2)) & 0x7c); }

struct item_pos {
	unsigned tile;   // k
	unsigned round;  // of `rounds`
	unsigned rounds;
	unsigned b0, b1; // the item's batches [b0, b1) of the step's spike list
};

// ticket `local` (0-based within its (connection, step)) -> tile, round and batch range
SPICE_PLAN_HD item_pos locate_item(unsigned local, bool arranged, unsigned total, unsigned per_round) {
	unsigned const nbatch = (total + 31) / 32;
	item_pos p;
	p.rounds           = rounds_of(arranged, total, per_round);
	unsigned const per = (nbatch + p.rounds - 1) / p.rounds; // <= per_round
	p.tile             = local / p.rounds;
	p.round            = local % p.rounds;
	p.b0               = p.round * per;
	p.b1               = p.b0 + per < nbatch ? p.b0 + per : nbatch;
	return p;
}
}
