// Internal interface of the tiled spike-delivery kernel (deliver.cu).
#pragma once

#include <cstdint>

#include "spice/detail/abi.h"

namespace spice::deliver {

constexpr int kTileMax = 5120; // targets per tile: a CTA keeps two u32 counters per target in shared memory (40 KB)

// One connection as the delivery kernel sees it (lives in device memory, one array per context,
// in schedule order: heaviest connections first).
struct conn_desc {
	// spike lists of the SOURCE population, one per ring slot.  tiles_args::world == 1: flat,
	// ids[slot * ring_cap .. + cnt[slot * cnt_stride]).  Several ranks: one segment per rank inside the slot, rank r's
	// spikes at ids[slot * ring_cap + seg_lo[r] ..], their number at cnt[slot * cnt_stride + r] (cnt_stride == world):
	// the spike ring as the update kernels and the peers' stores leave it.
	std::int32_t const* ring_ids;
	std::uint32_t const* ring_cnt;
	long long ring_cap;
	long long cnt_stride;
	std::int32_t seg_lo[spice::detail::kMaxWorld]; // first neuron of each rank's range in the source population
	std::int32_t const* packed;    // arranged connections: the delivery stream (see pack_runs)
	unsigned const* run_ptr;       // arranged connections: [src * tiles + 1] first 16-byte group of every run
	std::int32_t const* neighbors; // plain connections: CSR entries (local columns)
	long long const* tile_ptr;     // plain connections: [src][tiles + 1]: where each tile's share of the row starts
	std::uint32_t* counts;         // [cring][cstride] event counters of the TARGET population
	long long n_dst;               // local targets
	long long cstride;             // multiple of 8, >= n_dst
	long long delay;               // steps
	std::int32_t cring;
	std::int32_t tiles;            // number of target tiles
	std::int32_t tile;             // targets per tile (multiple of 256, <= kTileMax)
	std::int32_t tile_prefix;      // tiles of the connections scheduled before this one
	std::int32_t arranged;         // 1: packed stream (fast path); 0: plain columns, counted with
	                               //    global atomics (rows that may hold duplicate targets: adj_list multapses)
	std::int32_t pad;
};

struct tiles_args {
	conn_desc const* conns;
	int nconns;
	int total_tiles; // sum of conns[].tiles
	int ring, world;
	long long t0;
	int nsteps;
	unsigned* work;            // dynamic ticket counter, zeroed by the window prologue
	unsigned long long* stats; // [0] events, [1] spikes
	int* error;                // bit 16: internal error in the delivery kernel
	int tile_cap;              // targets per counter array (max conns[].tile rounded up to 128)
	// several ranks: the launch itself waits until every peer has published window `seq` (runtime.cu publish_window)
	unsigned long long const* flags; // this rank's flag array [world], or null
	unsigned long long seq;
	unsigned cnt_base;               // filled in by launch_tiles: where the CTA's counters start in its shared-memory window
	int rounds;                      // filled in by launch_tiles: parts of a unit's spike list handed out separately (1: whole units)
	int prezeroed;                   // the counters' consumer (the update kernel) clears what it has read: nothing is stored for a unit
	                                 // without spikes, and units may be counted in rounds that ADD to the counters
};

constexpr int kMaxConns  = 32;   // connections per launch
constexpr int kMaxCounts = 768;  // nconns * nsteps * world spike counts staged in shared memory per launch

// tile_ptr[row * (tiles + 1) + k] = first position in row `row` whose target is >= k * tile
// (k = tiles: the row end).  Returns a cudaError_t as int.
int build_tile_ptr(void* stream, long long const* offsets, std::int32_t const* neighbors, long long src_count, int tile,
                   int tiles, long long* tile_ptr);

// The fast path's own stream format, built once per duplicate-free connection from its CSR:
// every run (a tile's share of a row) becomes whole 16-byte groups of counter BYTE offsets (deliver_plan.h) —
// array A at word t, array B at word cap + rotw_fwd(t), a dump area of 32 words behind them for padding —
// permuted so that the 32 lanes of one counting instruction hit 32 different shared-memory banks.
// `cap` must equal tiles_args::tile_cap.
//   count_groups: run_ptr[src * tiles + 1] (exclusive scan of the runs' group counts) and their sum;
//   pack_runs:    fills `packed` (4 * groups entries);
//   unpack_rows:  the canonical (ascending, local column) CSR entries of [0, edges) into `out`.
int count_groups(void* stream, long long const* tile_ptr, long long src_count, int tiles, unsigned* run_ptr, long long* groups_out);
int pack_runs(void* stream, std::int32_t const* neighbors, long long const* tile_ptr, unsigned const* run_ptr, long long src_count,
              int tile, int tiles, int cap, std::int32_t* packed);
int unpack_rows(void* stream, std::int32_t const* packed, unsigned const* run_ptr, long long const* offsets, long long src_count,
                int tile, int tiles, int cap, std::int32_t* out);

// Delivers every spike of the window on every connection with a persistent grid filling the device; `launches`
// receives the number of kernels launched.  Returns a cudaError_t as int.
int launch_tiles(void* stream, tiles_args const& a, int device, int* launches = nullptr);

// Loads the delivery kernels (a lazily loaded kernel can synchronise the context at its first launch).
int preload();
}
