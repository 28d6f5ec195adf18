// Drop-in for the part of the reference's SoftwareRasterizer/VectorMath.h that application code uses
// (Main.cpp:96-113): the Aabb type handed to SurfaceAreaHeuristic::generateBatches and to Occluder::bake.
// Same member names and method set (VectorMath.h:26-66); the layout -- m_min then m_max, 32 bytes -- is
// exactly the (min4, max4) box record of the C ABI, so a std::vector<Aabb> is passed through as is.
#pragma once

#include <smmintrin.h>

#include <limits>

struct Aabb
{
	__m128 m_min = _mm_set1_ps(std::numeric_limits<float>::infinity());
	__m128 m_max = _mm_set1_ps(-std::numeric_limits<float>::infinity());

	// minps / maxps keep their second operand on ties: the grown box's zero bounds take the sign of the newcomer
	void include(__m128 point) { m_min = _mm_min_ps(m_min, point); m_max = _mm_max_ps(m_max, point); }
	void include(const Aabb& other) { m_min = _mm_min_ps(m_min, other.m_min); m_max = _mm_max_ps(m_max, other.m_max); }

	__m128 getCenter() const { return _mm_add_ps(m_min, m_max); }   // twice the centre, as in the reference
	__m128 getExtents() const { return _mm_sub_ps(m_max, m_min); }
	__m128 surfaceArea() const
	{
		const __m128 e = getExtents();
		return _mm_dp_ps(e, _mm_shuffle_ps(e, e, _MM_SHUFFLE(3, 0, 2, 1)), 0x7F);   // xy + yz + zx, half the area
	}
};
static_assert(sizeof(Aabb) == 32, "Aabb must stay two float4: it is the box record of orz_generate_batches");
