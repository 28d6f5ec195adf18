// Drop-in replacement for the reference's SoftwareRasterizer/Rasterizer.h (Rasterizer.h:10-61):
// same class name and public signatures; every method forwards to the C ABI in include/orz.h,
// which runs the CUDA kernels.  Per-call semantics as the reference: setModelViewProjection /
// clear / rasterize are asynchronous on the context's stream; queryVisibility / query2D return data
// to the host: boxes the frustum culls or the near plane clips are answered on the host, a rectangle
// test either finds its answer in the mapped mailbox (asked for by an earlier launch that predicted
// this query from the previous frame's sequence) or is launched and waited for;
// readBackDepth synchronises.
//
// Differences a caller can observe (both documented in DESIGN.md):
//   * clear() also zeroes the depth buffer ("fresh" state): the reference leaves stale depth
//     behind cleared blocks, which query2D then reads (Rasterizer.cpp:107-121 vs 310-343).
//   * many independent views (Main.cpp:181-206 per view) go through the C ABI's orz_render_views (include/orz.h);
//     this class keeps the reference's one-view interface.
#pragma once

#include <immintrin.h>

#include <cstdint>
#include <memory>
#include <vector>

struct Occluder;
struct orz_rasterizer;
struct orz_context;

class Rasterizer
{
public:
	Rasterizer(uint32_t width, uint32_t height);
	~Rasterizer();
	Rasterizer(const Rasterizer&) = delete;
	Rasterizer& operator=(const Rasterizer&) = delete;

	void setModelViewProjection(const float* matrix);

	void clear();

	template<bool possiblyNearClipped>
	void rasterize(const Occluder& occluder);

	bool queryVisibility(__m128 boundsMin, __m128 boundsMax, bool& needsClipping);

	bool query2D(uint32_t minX, uint32_t maxX, uint32_t minY, uint32_t maxY, uint32_t maxZ) const;

	void readBackDepth(void* target) const;

	// ---- additions (not in the reference) ----
	// many boxes in one call: out[i] bit0 = visible, bit1 = needsClipping
	void queryVisibilityBatch(const float* boxesMinMax, uint32_t count, uint8_t* out);
	// raw buffers in the reference layout: depth u16 [block][row][px], HiZ u16 [block]
	void download(uint16_t* depth, uint16_t* hiZ) const;
	static orz_context* context();   // the calling thread's context (created on first use, shared by the objects it creates)

private:
	orz_rasterizer* m_impl;
	uint32_t m_width;
	uint32_t m_height;
	std::shared_ptr<void> m_context;   // keeps the creating thread's context alive for as long as this object lives
};
