// Drop-in replacement for the reference's SoftwareRasterizer/QuadDecomposition.h (QuadDecomposition.h:6-10):
// same class and signature over orz_quad_decompose (include/orz.h), which returns the reference's quads
// in the reference's order.
#pragma once

#include <xmmintrin.h>

#include <cstdint>
#include <vector>

class QuadDecomposition
{
public:
	static std::vector<uint32_t> decompose(const std::vector<uint32_t>& indices, const std::vector<__m128>& vertices);
};
