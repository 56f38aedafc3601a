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
int generate_fixed_probability(void* cuda_stream, long long src, long long dst, double p, unsigned long long seed_lo,
                               unsigned long long seed_hi, long long col_lo, long long col_hi, long long chunk_draws,
                               result* out, std::string* err);
}
