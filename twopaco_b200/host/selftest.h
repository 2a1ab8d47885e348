// selftest.h -- `twopaco --test`: randomized end-to-end check of the GPU path against a
// brute-force junction finder (the role of the reference's src/graphconstructor/test.cpp).
#ifndef TWOPACO_B200_SELFTEST_H_
#define TWOPACO_B200_SELFTEST_H_

#include <cstddef>
#include <string>

namespace TwoPaCo
{
	// Runs `tests` random cases (one chromosome of `length` bp with sporadic 'N' + mutated copies)
	// for k = 3,5,7,9 and 1..4 rounds; true when every case matches the brute-force marks and every
	// brute-force junction has an id.  Same recipe as test.cpp:163-254 / constructor.cpp:147.
	bool RunTests(size_t tests, size_t filterBits, size_t length, size_t chrNumber, const std::string & temporaryDir);
}

#endif
