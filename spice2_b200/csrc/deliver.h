// Internal interface of the tiled spike-delivery kernel (deliver.cu).
#pragma once

#include <cstdint>

#include "spice/detail/abi.h"

namespace spice::deliver {

constexpr int kTileMax = 5120; // u16 counters one warp keeps in shared memory (10 KB)

// One connection as the delivery kernel sees it (lives in device memory, one array per context,
// in schedule order: heaviest connections first).
struct conn_desc {
	std::int32_t const* ring_ids;  // spike ring of the SOURCE population (this rank's copy)
	std::uint32_t const* ring_cnt; // [ring][world]
	long long ring_cap;
	long long seg_lo[spice::detail::kMaxWorld]; // first source neuron of every rank's segment of a ring slot
	std::int32_t const* neighbors; // CSR column indices (local to this rank's target range)
	long long const* tile_ptr;     // [src][tiles + 1]: where each tile's share of the row starts
	std::uint32_t* counts;         // [cring][cstride] event counters of the TARGET population
	long long n_dst;               // local targets
	long long cstride;             // multiple of 8, >= n_dst
	long long delay;               // steps
	std::int32_t cring;
	std::int32_t tiles;            // number of target tiles
	std::int32_t tile;             // targets per tile (multiple of 256, <= kTileMax)
	std::int32_t tile_prefix;      // tiles of the connections scheduled before this one
	std::int32_t atomic;           // rows may hold duplicate targets (adj_list multapses)
	std::int32_t pad;
};

struct tiles_args {
	conn_desc const* conns;
	int nconns;
	int total_tiles; // sum of conns[].tiles
	int ring, world;
	long long t0;
	int nsteps;
	unsigned* work;            // dynamic unit counter, zeroed by the window prologue
	unsigned long long* stats; // [0] events, [1] spikes
	int* error;                // bit 16: internal error in the delivery kernel
	int tile_cap;              // u16 counters per warp in shared memory (max conns[].tile)
};

// tile_ptr[row * (tiles + 1) + k] = first position in row `row` whose target is >= k * tile
// (k = tiles: the row end).  Returns a cudaError_t as int.
int build_tile_ptr(void* stream, long long const* offsets, std::int32_t const* neighbors, long long src_count, int tile,
                   int tiles, long long* tile_ptr);

// One launch delivers every spike of the window on every connection.  `blocks` <= 0 picks a
// persistent grid filling the device.  Returns a cudaError_t as int.
int launch_tiles(void* stream, tiles_args const& a, int device);
}
