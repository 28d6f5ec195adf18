// Drop-in replacement for the reference's SoftwareRasterizer/Occluder.h (Occluder.h:7-21):
// same struct name, same static bake() signature, same public data members, so application code
// written against the reference (Main.cpp:116-128, 186-195) compiles unchanged.  bake() runs on
// the host through the C ABI (orz_bake); the device copy is created lazily by the first
// Rasterizer::rasterize call and released by the destructor the reference never had.
#pragma once

#include <immintrin.h>

#include <cstdint>
#include <memory>
#include <vector>

struct orz_occluder;

struct Occluder
{
	static std::unique_ptr<Occluder> bake(const std::vector<__m128>& vertices, __m128 refMin, __m128 refMax);

	Occluder() = default;
	Occluder(const Occluder&) = delete;
	Occluder& operator=(const Occluder&) = delete;
	~Occluder();

	__m128 m_center;

	__m128 m_refMin;
	__m128 m_refMax;

	__m128 m_boundsMin;
	__m128 m_boundsMax;

	__m256i* m_vertexData = nullptr;   // reference packet layout (Occluder.cpp:146-156), 32-byte aligned
	uint32_t m_packetCount = 0;

	mutable orz_occluder* m_device = nullptr;   // HBM-resident copy (one 16-byte record per quad)
	mutable std::shared_ptr<void> m_context;    // context that owns m_device, kept alive until the destructor has run
};
