// Precondition / invariant checks with the reference's error convention: a violated condition
// throws std::logic_error("Assertion failed (<file>:<line>): <condition>")
// (reference: spice/util/assert.h:3-17, spice/src/util/assert.cpp:8-15).  Host only.
#pragma once

#include <stdexcept>
#include <string>

namespace spice::util::detail {
[[noreturn]] inline void assert_failed(char const* file, int line, char const* condition) {
	std::string f(file);
	auto const slash = f.find_last_of('/');
	if (slash != std::string::npos)
		f = f.substr(slash + 1);
	throw std::logic_error("Assertion failed (" + f + ":" + std::to_string(line) + "): " + condition);
}
}

#define SPICE_ASSERT(X)                \
	do {                               \
		if (__builtin_expect(!(X), 0)) \
			::spice::util::detail::assert_failed(__FILE__, __LINE__, #X); \
	} while (0)

// Both classes of check are always on in this backend (the reference's release build enables
// them too, CMakeLists.txt:6-7): they run on the host, outside the kernels.
#define SPICE_PRE(X) SPICE_ASSERT(X)
#define SPICE_INV(X) SPICE_ASSERT(X)
// in a function that also compiles for the device: checked where it can throw
#ifdef __CUDA_ARCH__
#define SPICE_PRE_HOST(X) ((void)0)
#else
#define SPICE_PRE_HOST(X) SPICE_ASSERT(X)
#endif
