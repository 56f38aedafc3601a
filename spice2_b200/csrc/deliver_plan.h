// Work items of a split delivery launch (deliver.cu, deliver_tiles<true>): pure arithmetic, shared by the
// planning kernel, the delivery kernel and the host test (tests/test_host.py::test_delivery_plan_covers_every_batch).
//
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
