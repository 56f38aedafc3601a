// Internal C++ interface of the fixed_probability generator (generator.cu).
#pragma once

#include <string>

namespace spice::gen {
struct result {
	long long* offsets = nullptr; // device, int64[src + 1]
	int* neighbors     = nullptr; // device, int32[edges] (local column indices)
	long long edges    = 0;
	long long draws    = 0;       // length of the consumed engine stream
	long long exact_rows = 0;     // rows whose end needed the exact replay
	float total_ms     = 0;
	float rows_ms      = 0;       // fp_rows kernel time (sum over chunks)
	int launches       = 0;
};

// trunc(dst*p + 3*sqrt(dst*p*(1-p))) with the reference build's fused form (topology.cpp:75-78)
long long max_degree(long long dst, double p);

// Returns 0 or a SPICE_ERR_* code (message in *err).  chunk_draws <= 0 picks the default.
// adj_list::generate on the GPU (topology.cpp:56-71: sort the packed (src << 32 | dst) connections, stream them into
// CSR): radix sort of the packed keys, rows by a histogram + scan.  edges_src / edges_dst are HOST arrays of n_edges
// entries (what adj_list::connect collected); targets in [col_lo, col_hi) are kept as local columns.  *duplicates = a
// (src, dst) pair occurs more than once (a multapse).  Returns 0, SPICE_ERR_PRECONDITION (1: an index out of range,
// the reference's SPICE_PRE in edge_stream) or a CUDA error code as generate_fixed_probability does.
int generate_adj_list(void* cuda_stream, int const* edges_src, int const* edges_dst, long long n_edges, long long src, long long dst,
                      long long col_lo, long long col_hi, result* out, bool* duplicates, std::string* err);

// The counter-based generator (one engine per (row, lane): seed_seq::stream): independent Bernoulli(p) per pair by geometric
// skips, rows in parallel, bounded by write bandwidth.  Same interface; NOT the reference's matrix for the same seed.
int generate_fixed_probability_fast(void* cuda_stream, long long src, long long dst, double p, unsigned long long seed_lo,
                                    unsigned long long seed_hi, long long col_lo, long long col_hi, result* out, std::string* err);

int generate_fixed_probability(void* cuda_stream, long long src, long long dst, double p, unsigned long long seed_lo,
                               unsigned long long seed_hi, long long col_lo, long long col_hi, long long chunk_draws,
                               result* out, std::string* err);
}
