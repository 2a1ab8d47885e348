// vertexenumerator.h -- host-side mirror of the reference's operator interface for the
// junction-finding path: TwoPaCo::VertexEnumerator / TwoPaCo::CreateEnumerator
// (reference: src/graphconstructor/vertexenumerator.h:23-46).  Same names, argument order and
// meaning, same error behaviour (std::runtime_error with the reference's messages), but the
// work is done by libtwopaco_b200.so on the current CUDA device through the C ABI
// (include/twopaco_b200.h).  A maintainer of the reference can drop this header + the library
// in place of vertexenumerator.{h,cpp}: constructor.cpp and test.cpp compile unchanged against
// it (GetHashSeed() is the one member not provided: hashing lives on the GPU and its seed is
// unobservable in the output).
#ifndef TWOPACO_B200_VERTEX_ENUMERATOR_H_
#define TWOPACO_B200_VERTEX_ENUMERATOR_H_

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <numeric>
#include <ostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

// The reference's header also makes its src/common helpers visible to its includers (constructor.cpp, test.cpp use
// DnaChar and JunctionPositionReader through it, vertexenumerator.h:12-19); inside the reference tree those files stay.
#if defined(__has_include)
#if __has_include(<junctionapi.h>)
#include <junctionapi.h>
#endif
#if __has_include(<dnachar.h>)
#include <dnachar.h>
#endif
#endif

#if defined(__has_include) && __has_include("twopaco_b200.h")
#include "twopaco_b200.h"
#else
#include "../../include/twopaco_b200.h"
#endif

namespace TwoPaCo
{
	const int64_t INVALID_VERTEX = TPC_INVALID_VERTEX;  // common.cpp:5

	class VertexEnumerator
	{
	public:
		virtual size_t GetVerticesCount() const = 0;
		virtual int64_t GetId(const std::string & vertex) const = 0;
		virtual ~VertexEnumerator() {}
	};

	std::unique_ptr<VertexEnumerator> CreateEnumerator(const std::vector<std::string> & fileName,
		size_t vertexLength,
		size_t filterSize,
		size_t hashFunctions,
		size_t rounds,
		size_t threads,
		size_t abundance,
		const std::string & tmpFileName,
		const std::string & outFileName,
		std::ostream & logStream);
}

#endif
